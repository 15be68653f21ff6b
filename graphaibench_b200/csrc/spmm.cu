// Neighbour aggregation (CSR SpMM) for sm_100a: GCN symmetric-normalised, SAGE mean (forward / transposed) and
// explicit edge values (GAT), one kernel family.
//
// Replaces update_all_gcn / update_all_sage / reduce_warp+reduce_cta (include/gnn/graph_operations.h:8-178) and the
// CPU loops they mirror (src/gnn/gconv/gcn_aggregator.cpp:48-77, sage_aggregator.cpp:7-54, gat_aggregator.cpp:26-45).
//
// Design (B200: an HBM/L2 gather; no tensor-core shape here):
//   * every input row is read with 128-bit loads. Rows whose pitch is a multiple of 4 floats are read in place; only a
//     caller's dense matrix with F % 4 != 0 (or a misaligned base) goes through a zero-padded staging copy.
//     What bounds the kernel (ncu + tools/l2_probe.cu, profiles/README.md round 2): the LATENCY of the gathers — 8.7 TB/s of L2 -> SM
//     traffic for F = 100 against a measured 19 TB/s L2 streaming rate, DRAM at 49 %, occupancy 48 % at 64 registers, long_scoreboard the
//     top stall: the bytes in flight per SM are capped by the registers that hold them. It is NOT the L1 data pipe: 128-byte-aligned row
//     pitches cut its wavefronts by a third and changed nothing (the larger footprint cost the hit rates), predicated lanes likewise.
//     Landing the gathers in shared memory instead of registers (a private 16-slot cp.async FIFO per lane, 8..16 row chunks in flight
//     per lane at all times, no barrier or second warp role; bit-exact, round 2) was measured at 5.05 ms against 2.81 ms for F = 100:
//     every gathered byte then crosses the shared-memory pipe twice (LDGSTS write + LDS read, 128 B/clk/SM), which costs more than the
//     extra bytes in flight buy. Together with the earlier per-row cp.async.bulk and producer/consumer LDGSTS-ring attempts this rules
//     out shared-memory staging for the light rows: the register file is the cheapest landing zone this gather has.
//   * one persistent kernel (4 CTAs x 148 SMs), two kinds of work items taken from global counters:
//       light rows (deg <= hub_degree): rows in DEGREE order, cut into claims of <= 32 rows / <= 2048 edges. A group of
//         G lanes (G = 4..32, from the feature width) owns one output row in registers; the group loads G column
//         indices with one coalesced request, broadcasts them by shuffle, and keeps U independent 128-bit
//         neighbour-row loads in flight per lane, with a two-deep index pipeline across rows and batches.
//       hub rows (deg > hub_degree, a per-graph threshold = a warp's fair share of the edges): item = (row, column block
//         of <= 4 float4 chunks). A whole CTA takes one item: 7 producer warps fetch (cp.async) + scale stages of 64 edges into a
//         shared-memory ring (mbarrier full/empty pairs); warp 0 adds the staged products IN EDGE ORDER. The longest
//         rows are listed first, so their add chains run underneath the light rows.
//       (widths beyond 512 floats keep the older CTA-per-row kernel spmm_hub_kernel, forked onto a side stream.)
//   * numerics: acc = fadd_rn(acc, fmul_rn(w, x)) per edge, sequential per column — exactly the reference CPU path's
//     scale()+vadd() (math_functions.cpp:266,336): results are bit-identical for every row length, hub rows included.
//   * fused: zero-init (no memset pass), optional "+ addend", ReLU and sign-bit d_relu mask epilogues, leading
//     dimensions, row ranges (1D partition: interior vs boundary rows).
#include <cstdlib>
#include "gai_internal.cuh"

namespace {

// M_EDGE_H / M_EDGE_PERM_H: multi-head edge values, vals[e * heads + h] with h = the head the output column belongs to (GAT extension)
enum Mode { M_GCN = 0, M_MEAN = 1, M_MEAN_T = 2, M_EDGE = 3, M_EDGE_PERM = 4, M_EDGE_H = 5, M_EDGE_PERM_H = 6 };

struct SpmmArgs {
  const uint32_t* rowptr;
  const uint32_t* colidx;
  const float* norm;
  const float* vals;
  const uint32_t* perm;
  const float* in;   // rows 16-byte aligned, ld_in % 4 == 0
  float* out;
  const float* addend;
  int F;             // logical width (columns written)
  int nchunks;       // ceil(F / 4): float4 chunks read per neighbour row
  int ld_in, ld_out;
  uint32_t row_begin, row_end;
  int mode, flags;
  int out_vec;       // 1: out/addend rows are 16-byte aligned and F % 4 == 0 -> float4 epilogue
  uint32_t hub_threshold;
  int hub_per;       // hub rows: float4 chunks per CTA column block (gridDim.y blocks cover nchunks)
  const uint32_t* mask_bits;  // optional d_relu of the layer below: out = bit ? out : 0, one sign bit per element (GAI_EPI_BITMASK)
  int ld_bits;                // words per row
  // 1D partition: neighbour ids >= n_split are halo vertices, whose rows live in a separate buffer (row = id - n_split, same pitch):
  // the owners' rows land there (gai_halo_pull) and the matrix itself carries master rows only. n_split = 0xffffffff: one matrix.
  const float* in_halo;
  uint32_t n_split;
  int hub_cl;          // float4 chunks per hub column block (8 or 4): launch_rows_mode
  int heads, hshift;   // multi-head edge values: head of float4 chunk c = c >> hshift (columns per head = 4 << hshift)
};

// base of neighbour row `c` (float4 units). SPLIT is a template parameter of the kernels: the single-matrix path pays nothing for it
// (a run-time test per gathered row cost 8 % of the aggregation time).
// Address of a 16-byte chunk of neighbour row `c`: base + c * row_bytes with a 32-bit row pitch, i.e. ONE IMAD.WIDE.U32 per gathered row
// (`base` already carries the lane's chunk offset). The aggregation is issue-bound enough (70 % of the issue slots at F = 100) for the
// address arithmetic to matter: a 64-bit pitch cost 5 instructions per edge (13.75 in all), a second address computation for the halo
// block cost 30 % (N = 8: 5.7 ms vs 4.2 ms for the same rows in one matrix).
// `halo` is the VIRTUAL base of the halo block (halo_virtual_base: first halo row minus n_split rows), so that both branches of a split
// (masters | halo) input share the multiply-add and differ in the base pointer only.
template <bool SPLIT>
__device__ __forceinline__ const float4* row_chunk(const char* base, const char* halo, uint32_t n_split, uint32_t row_bytes, uint32_t c) {
  const char* b = (SPLIT && c >= n_split) ? halo : base;
  return reinterpret_cast<const float4*>(b + (unsigned long long)c * row_bytes);
}
__device__ __forceinline__ const char* halo_virtual_base(const float* in_halo, uint32_t n_split, uint32_t row_bytes) {
  return reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(in_halo) - (uintptr_t)n_split * row_bytes);
}

// out[row, 4*chunk .. 4*chunk+3] = epilogue(acc)
__device__ __forceinline__ void store_chunk(const SpmmArgs& a, uint32_t row, int chunk, float4 r) {
  const size_t o = (size_t)row * a.ld_out + (size_t)chunk * 4;
  if (a.out_vec) {
    if (a.flags & GAI_EPI_ADD) {
      const float4 ad = *reinterpret_cast<const float4*>(a.addend + o);
      r.x = __fadd_rn(r.x, ad.x); r.y = __fadd_rn(r.y, ad.y); r.z = __fadd_rn(r.z, ad.z); r.w = __fadd_rn(r.w, ad.w);
    }
    if (a.flags & GAI_EPI_RELU) { r.x = r.x > 0.f ? r.x : 0.f; r.y = r.y > 0.f ? r.y : 0.f; r.z = r.z > 0.f ? r.z : 0.f; r.w = r.w > 0.f ? r.w : 0.f; }
    if (a.mask_bits) {
      const uint32_t nib = __ldg(a.mask_bits + (size_t)row * a.ld_bits + (chunk >> 3)) >> ((chunk & 7) * 4);
      r.x = (nib & 1u) ? r.x : 0.f; r.y = (nib & 2u) ? r.y : 0.f; r.z = (nib & 4u) ? r.z : 0.f; r.w = (nib & 8u) ? r.w : 0.f;
    }
    *reinterpret_cast<float4*>(a.out + o) = r;
  } else {
    const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (chunk * 4 + k < a.F) {
        float t = v[k];
        if (a.flags & GAI_EPI_ADD) t = __fadd_rn(t, a.addend[o + k]);
        if (a.flags & GAI_EPI_RELU) t = t > 0.f ? t : 0.f;
        if (a.mask_bits) { const int c = chunk * 4 + k; if (!((__ldg(a.mask_bits + (size_t)row * a.ld_bits + (c >> 5)) >> (c & 31)) & 1u)) t = 0.f; }
        a.out[o + k] = t;
      }
    }
  }
}

// ---- light rows -----------------------------------------------------------------------------------------------------
// One output row per group of G lanes, kept in registers; a warp owns 32/G rows at a time.
//
// Work list: rows in DEGREE order (g->row_order, longest first, built once per graph), hub rows excluded. The rows a
// warp processes together therefore have (nearly) equal lengths — no lane group idles while its sibling finishes a long
// row — the longest rows start first, and the empty rows (40% of an R-MAT graph) end up in all-empty warps.
// A warp claims 32 consecutive list entries with one atomic, loads their row ids and row bounds with one coalesced
// request each into shared memory, and then walks them with a two-deep index pipeline:
//     column indices of the NEXT row's first batch and of the NEXT batch of this row are requested before the current
//     batch's neighbour rows are gathered, so a row costs one exposed memory latency (the gather) instead of three
//     (rowptr -> colidx -> gather).
// Arithmetic per edge and column: p = mul.rn(w, x); acc = add.rn(acc, p) — the reference's scale()+vadd(). The multiply
// is issued as FMUL2 (two columns per instruction); the add stays scalar: ptxas contracts a packed mul + packed add
// pair into FFMA2 even with .rn on both (12.9), which would break bit-exactness.
__device__ __forceinline__ void mul_w_f4(float w, const float4& x, float4& p) {
  unsigned long long w2, lo, hi;
  asm("mov.b64 %0, {%1, %1};" : "=l"(w2) : "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(lo) : "f"(x.x), "f"(x.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(hi) : "f"(x.z), "f"(x.w));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(lo) : "l"(w2), "l"(lo));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(hi) : "l"(w2), "l"(hi));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(p.x), "=f"(p.y) : "l"(lo));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(p.z), "=f"(p.w) : "l"(hi));
}
__device__ __forceinline__ void acc_add(float4& acc, const float4& p) {
  acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y); acc.z = __fadd_rn(acc.z, p.z); acc.w = __fadd_rn(acc.w, p.w);
}

// 128-bit gather of a neighbour-row chunk. GAI_GATHER_NOALLOC (build-time experiment): read-only path without L1 allocation.
__device__ __forceinline__ float4 gather4(const float4* p) {
#ifdef GAI_GATHER_NOALLOC
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}

template <int MODE>
__device__ __forceinline__ float edge_weight_t(const SpmmArgs& a, float wrow, uint32_t idx, uint32_t c) {
  if (MODE == M_GCN) return __fmul_rn(wrow, __ldg(a.norm + c));  // b = a_i * a_j (gcn_aggregator.cpp:66)
  if (MODE == M_MEAN) return wrow;                               // 1/deg_i (sage_aggregator.cpp:17)
  if (MODE == M_MEAN_T) return __ldg(a.norm + c);                // 1/deg_j (sage_aggregator.cpp:41)
  if (MODE == M_EDGE) return __ldg(a.vals + idx);
  if (MODE == M_EDGE_H) return 0.0f;                                               // weights are fetched per (edge, head) at the gather
  if (MODE == M_EDGE_PERM_H) return __uint_as_float(__ldg(a.perm + idx));          // the reverse edge's index, carried as bits
  return __ldg(a.vals + __ldg(a.perm + idx));
}
template <int MODE>
__host__ __device__ constexpr bool mode_has_heads() { return MODE == M_EDGE_H || MODE == M_EDGE_PERM_H; }

// ---- hub rows inside the persistent light-row kernel --------------------------------------------------------------------
// A hub item = (hub row, column block of <= 4 float4 chunks). A whole 256-thread CTA of the persistent kernel takes one item at a
// time before it turns to the light-row claims: warp 0 adds in edge order (one lane per chunk), warps 1..7 fetch + scale stages
// of 64 edges into a 7-slot shared-memory ring (mbarrier full/empty pairs), eight edges per cp.async instruction so that every lane
// carries a 128-bit request. The longest rows come first in the item list, so their sequential add chains start at t = 0 and run
// underneath the rest of the kernel; no second kernel, stream or event is involved and no SM is ever reserved for hub rows.
// A hub row is a latency-bound pipeline — ring bytes / (memory round trip of a stage) — and on partitioned graphs, whose rows reach
// 10^5..10^6 edges, the longest row IS the kernel (tools/shard_probe.py, GAI_SPMM_DEBUG=2: 3.1 ms of a 3.3 ms call). What counts is
// therefore EDGES in flight per ring byte: column blocks of 4 chunks (64 bytes per neighbour row) put 448 edges in flight per CTA
// where blocks of 8 chunks put 224, at the price of twice as many items.
// The block width is chosen per call (hub_block_chunks): 8 chunks while the longest row stays below 128 K edges (fewer items: on one GPU
// the hub chains hide under the light rows and narrower blocks only add items, +3..5 %), 4 chunks beyond (partitioned graphs).
constexpr int HI_PW = 7;            // producer warps = ring slots
constexpr int HI_SLOT_F4 = 256;     // float4 entries per ring slot = (edges per stage) x (chunks per column block): 32 x 8 or 64 x 4
constexpr int HI_RING_F4 = HI_PW * HI_SLOT_F4;   // float4 entries of the ring (28 KB)
struct HubShared {
  uint64_t full_bar[HI_PW], empty_bar[HI_PW];
  uint32_t round0[HI_PW];  // uses of each slot by the items this CTA has already processed (mbarrier phase bookkeeping)
  unsigned long long item;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}


__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int MODE, bool SPLIT, int HI_CL>
__device__ __forceinline__ void hub_item_cta(const SpmmArgs& a, uint32_t row, uint32_t s, uint32_t e, int cb, int nch, HubShared& sh,
                                             float4* __restrict__ ring /* in a register: read through `sh` it cost the in-order consumer a
                                             shared-memory load per stage, and on partitioned graphs (rows of 10^5..10^6 edges) that add
                                             chain is the kernel's critical path: +70 % on the F = 47 calls at N = 4 */) {
  constexpr int HI_ES = HI_SLOT_F4 / HI_CL;   // edges per stage
  constexpr int NR = HI_ES / 32;              // index registers per producer lane
  const uint32_t nstages = (e - s + HI_ES - 1) / HI_ES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)a.ld_in * 4u;
  const char* inb = reinterpret_cast<const char*>(a.in);
  const char* halob = halo_virtual_base(a.in_halo, a.n_split, row_bytes);
  if (warp > 0) {
    // ---------------- producers ----------------
    const int pw = warp - 1;
    const float wrow = (MODE == M_GCN || MODE == M_MEAN) ? __ldg(a.norm + row) : 0.0f;
    float4* slot = ring + pw * HI_SLOT_F4;
    const uint32_t round0 = sh.round0[pw];
    // lane l holds the column index / weight of edges l, l + 32, ... of the stage
    uint32_t c[NR]; float w[NR];
#pragma unroll
    for (int h = 0; h < NR; h++) {
      c[h] = 0; w[h] = 0.0f;
      const uint64_t idx = (uint64_t)s + (uint64_t)pw * HI_ES + 32 * h + lane;
      if (idx < e) { c[h] = __ldg(a.colidx + idx); w[h] = edge_weight_t<MODE>(a, wrow, (uint32_t)idx, c[h]); }
    }
    constexpr int EL = 32 / HI_CL;
    const int cl = lane % HI_CL, el = lane / HI_CL;
    const bool chv = cl < nch;
    const int hd = (cb + cl) >> a.hshift;   // multi-head modes: the head of this lane's chunk
    const char* inb_c = inb + (size_t)(cb + cl) * 16;
    const char* halob_c = halob + (size_t)(cb + cl) * 16;
    asm volatile("" : "+l"(inb_c), "+l"(halob_c));
    for (uint32_t k = pw, r = 0; k < nstages; k += HI_PW, r++) {
      const uint32_t base = s + k * HI_ES;
      const int cnt = (e - base) < (uint32_t)HI_ES ? (int)(e - base) : HI_ES;
      uint32_t cur_c[NR]; float cur_w[NR];
#pragma unroll
      for (int h = 0; h < NR; h++) { cur_c[h] = c[h]; cur_w[h] = w[h]; }
#pragma unroll
      for (int h = 0; h < NR; h++) {
        c[h] = 0; w[h] = 0.0f;
        const uint64_t nidx = (uint64_t)base + (uint64_t)HI_PW * HI_ES + 32 * h + lane;
        if (nidx < e) { c[h] = __ldg(a.colidx + nidx); w[h] = edge_weight_t<MODE>(a, wrow, (uint32_t)nidx, c[h]); }
      }
      mbar_wait(&sh.empty_bar[pw], ((round0 + r) & 1) ^ 1);
      // The neighbour-row chunks of the whole stage are requested at once and land straight in the ring slot (cp.async: no register holds
      // them), then every lane scales the entries it requested itself, in place: one memory round trip per stage.
      constexpr int NI = HI_ES / EL;
      float ww[NI];
#pragma unroll
      for (int u = 0; u < NI; u++) {
        const int j = u * EL + el;
        const uint32_t cc = __shfl_sync(0xffffffffu, cur_c[(u * EL) >> 5], j & 31);
        ww[u] = __shfl_sync(0xffffffffu, cur_w[(u * EL) >> 5], j & 31);
        if (chv && j < cnt) cp_async16(smem_u32(slot + j * HI_CL + cl), row_chunk<SPLIT>(inb_c, halob_c, a.n_split, row_bytes, cc));
        if (mode_has_heads<MODE>() && chv && j < cnt) {
          const uint32_t eidx = MODE == M_EDGE_H ? base + (uint32_t)j : __float_as_uint(ww[u]);
          ww[u] = __ldg(a.vals + (size_t)eidx * a.heads + hd);
        }
      }
      cp_async_commit();
      cp_async_wait_all();
#pragma unroll
      for (int u = 0; u < NI; u++) {
        const int j = u * EL + el;
        if (chv && j < cnt) {
          const float4 x = slot[j * HI_CL + cl];
          float4 p;
          p.x = __fmul_rn(ww[u], x.x); p.y = __fmul_rn(ww[u], x.y); p.z = __fmul_rn(ww[u], x.z); p.w = __fmul_rn(ww[u], x.w);
          slot[j * HI_CL + cl] = p;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.full_bar[pw]);
    }
  } else {
    // ---------------- consumer: in-order add, one float4 chunk per lane ----------------
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t k = 0; k < nstages; k++) {
      const int pw = k % HI_PW;
      const uint32_t r = k / HI_PW;
      const uint32_t base = s + k * HI_ES;
      const int cnt = (e - base) < (uint32_t)HI_ES ? (int)(e - base) : HI_ES;
      mbar_wait(&sh.full_bar[pw], (sh.round0[pw] + r) & 1);
      if (lane < nch) {
        const float4* tile = ring + pw * HI_SLOT_F4 + lane;
        if (cnt == HI_ES) {
          constexpr int SB = 8;
          float4 p[2][SB];
#pragma unroll
          for (int j = 0; j < SB; j++) p[0][j] = tile[j * HI_CL];
#pragma unroll
          for (int b = 0; b < HI_ES / SB; b++) {
            if (b + 1 < HI_ES / SB) {
#pragma unroll
              for (int j = 0; j < SB; j++) p[(b + 1) & 1][j] = tile[((b + 1) * SB + j) * HI_CL];
            }
#pragma unroll
            for (int j = 0; j < SB; j++) acc_add(acc, p[b & 1][j]);
          }
        } else {
          for (int j = 0; j < cnt; j++) acc_add(acc, tile[j * HI_CL]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty_bar[pw]);
    }
    if (lane < nch) store_chunk(a, row, cb + lane, acc);
  }
}

// All hub items, one at a time per CTA (work counter counter[1]); returns when the list is exhausted.
template <int MODE, bool SPLIT>
__device__ __forceinline__ void hub_phase(const SpmmArgs& a, HubShared& hub_sh, float4* ring, unsigned long long* __restrict__ counter,
                                          const uint32_t* __restrict__ hub_rows, unsigned long long n_hub_items, int hub_nsplit) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < HI_PW; i++) { mbar_init(&hub_sh.full_bar[i], 1); mbar_init(&hub_sh.empty_bar[i], 1); hub_sh.round0[i] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (;;) {
    __syncthreads();  // previous item fully drained (ring, round0) / barriers initialised
    if (threadIdx.x == 0) hub_sh.item = atomicAdd(counter + 1, 1ull);
    __syncthreads();
    const unsigned long long item = hub_sh.item;
    if (item >= n_hub_items) break;  // uniform
    const uint32_t hrow = __ldg(hub_rows + item / (unsigned)hub_nsplit);
    const int cb = (int)(item % (unsigned)hub_nsplit) * a.hub_per;
    const int nch = (a.nchunks - cb) < a.hub_per ? (a.nchunks - cb) : a.hub_per;
    if (hrow < a.row_begin || hrow >= a.row_end || nch <= 0) continue;  // uniform
    const uint32_t hs = __ldg(a.rowptr + hrow), he = __ldg(a.rowptr + hrow + 1);
    if (a.hub_cl == 4) hub_item_cta<MODE, SPLIT, 4>(a, hrow, hs, he, cb, nch, hub_sh, ring);
    else hub_item_cta<MODE, SPLIT, 8>(a, hrow, hs, he, cb, nch, hub_sh, ring);
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t es = (uint32_t)(HI_SLOT_F4 / a.hub_cl);
      const uint32_t nst = (he - hs + es - 1) / es;
      for (int i = 0; i < HI_PW; i++) hub_sh.round0[i] += (nst + HI_PW - 1 - i) / HI_PW;
    }
  }
  __syncthreads();   // the ring may be reused by the caller
}

constexpr int SLOTS = 32;  // work-list entries per claim

template <int MODE, int G, int K, bool SPLIT>
__global__ void __launch_bounds__(256, MODE == M_EDGE_PERM_H ? 3 : 4) spmm_rows_kernel(  // the transposed multi-head mode carries 8 more registers of weights
const SpmmArgs a, const uint32_t* __restrict__ order, const uint32_t* __restrict__ claim_ptr,
                                                         unsigned long long n_claims, unsigned long long* __restrict__ counter,
                                                         const uint32_t* __restrict__ hub_rows, unsigned long long n_hub_items, int hub_nsplit) {
  constexpr int RPW = 32 / G;                                  // rows in flight per warp
  constexpr int UMAX = (K == 1) ? 8 : (K == 2 ? 4 : 2);
  constexpr int U = G < UMAX ? G : UMAX;                       // independent neighbour rows in flight per lane
  __shared__ uint32_t sm_row[8][SLOTS], sm_s[8][SLOTS], sm_e[8][SLOTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gl = lane % G, grp = lane / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const uint32_t row_bytes = (uint32_t)a.ld_in * 4u;
  const char* inb = reinterpret_cast<const char*>(a.in);
  const char* halob = halo_virtual_base(a.in_halo, a.n_split, row_bytes);
  // lanes whose chunk lies past the row width re-read chunk 0 (same sectors as lane 0) and store nothing. Predicating them off instead
  // was measured slower (round 2: F = 47 calls +4..9 %): the L1 data pipe is not the limiter and the predicate costs issue slots.
  int chunk[K];
  bool act[K];
#pragma unroll
  for (int k = 0; k < K; k++) { act[k] = (gl + G * k) < a.nchunks; chunk[k] = act[k] ? gl + G * k : 0; }

  // ---- hub items first (counter[1]), CTA-wide ----
  if (n_hub_items != 0) {
    __shared__ HubShared hub_sh;
    __shared__ float4 hub_ring[HI_RING_F4];
    hub_phase<MODE, SPLIT>(a, hub_sh, hub_ring, counter, hub_rows, n_hub_items, hub_nsplit);
  }

  for (;;) {
    unsigned long long claim = 0;
    if (lane == 0) claim = atomicAdd(counter, 1ull);
    claim = __shfl_sync(0xffffffffu, claim, 0);
    if (claim >= n_claims) break;
    {
      // ordered list: claim boundaries come from claim_ptr; natural order (row ranges): 32 consecutive rows per claim
      unsigned long long base, end;
      if (order) { base = __ldg(claim_ptr + 2 * claim); end = __ldg(claim_ptr + 2 * claim + 1); }
      else { base = claim * SLOTS; end = base + SLOTS; const unsigned long long n = (unsigned long long)a.row_end - a.row_begin; if (end > n) end = n; }
      uint32_t r = 0xffffffffu, s = 0, e = 0;
      if (base + lane < end) {
        r = order ? __ldg(order + base + lane) : (uint32_t)(a.row_begin + base + lane);
        if (r >= a.row_begin && r < a.row_end) { s = __ldg(a.rowptr + r); e = __ldg(a.rowptr + r + 1); } else r = 0xffffffffu;
        if (e - s > a.hub_threshold) r = 0xffffffffu;  // hub rows belong to the CTA-per-row kernel
      }
      __syncwarp();
      sm_row[warp][lane] = r; sm_s[warp][lane] = s; sm_e[warp][lane] = e;
      __syncwarp();
    }
    // first batch of the first row of this group
    uint32_t c_first = 0;
    {
      const uint32_t s0 = sm_s[warp][grp], e0 = sm_e[warp][grp];
      if (sm_row[warp][grp] != 0xffffffffu && s0 + gl < e0) c_first = __ldg(a.colidx + s0 + gl);
    }
#pragma unroll 1
    for (int it = 0; it < SLOTS / RPW; it++) {
      const int slot = it * RPW + grp;
      const uint32_t row = sm_row[warp][slot];
      const uint32_t s = sm_s[warp][slot], e = sm_e[warp][slot];
      // request the first batch of the next row this group will process
      uint32_t c_nextrow = 0;
      if (it + 1 < SLOTS / RPW) {
        const uint32_t ns = sm_s[warp][slot + RPW], ne = sm_e[warp][slot + RPW];
        if (sm_row[warp][slot + RPW] != 0xffffffffu && ns + gl < ne) c_nextrow = __ldg(a.colidx + ns + gl);
      }
      if (row != 0xffffffffu) {
        const float wrow = (MODE == M_GCN || MODE == M_MEAN) ? __ldg(a.norm + row) : 0.0f;
        for (int cb = 0; cb < a.nchunks; cb += G * K) {  // one pass unless F > 128*K
          float4 acc[K];
#pragma unroll
          for (int k = 0; k < K; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          int ch[K];
          bool av[K];
#pragma unroll
          for (int k = 0; k < K; k++) { av[k] = cb == 0 ? act[k] : (cb + gl + G * k) < a.nchunks; ch[k] = cb == 0 ? chunk[k] : (av[k] ? cb + gl + G * k : 0); }
          int hd[K];           // multi-head modes: the head of the lane's chunk k
#pragma unroll
          for (int k = 0; k < K; k++) hd[k] = ch[k] >> a.hshift;
          const char* bk[K];   // lane's chunk k of row 0 (masters) / of virtual row 0 (halo)
          const char* hk[K];
#pragma unroll
          for (int k = 0; k < K; k++) {
            bk[k] = inb + (size_t)ch[k] * 16; hk[k] = halob + (size_t)ch[k] * 16;
            asm volatile("" : "+l"(bk[k]), "+l"(hk[k]));  // keep base + chunk offset as ONE 64-bit addend (ptxas otherwise re-associates: +2 instructions per edge)
          }
          uint32_t c_cur = c_first;
          if (cb != 0 && s + gl < e) c_cur = __ldg(a.colidx + s + gl);
          for (uint32_t b = s; b < e; b += G) {
            const uint32_t idx = b + gl;
            uint32_t c_nb = 0;
            if (idx + G < e) c_nb = __ldg(a.colidx + idx + G);  // next batch of this row
            const float w = idx < e ? edge_weight_t<MODE>(a, wrow, idx, c_cur) : 0.0f;
            const int cnt = (e - b) < (uint32_t)G ? (int)(e - b) : G;
            // Multi-head modes with 8 heads (the configs[2] shape): the G x 8 weights of the batch live in 8 registers per lane, register r of
            // group lane l holding flat entry r * G + l = (edge, head) = ((r * G + l) / 8, l % 8). They are loaded with 8 requests per
            // BATCH (coalesced when the edge values are in edge order, 32-byte runs through the transpose permutation) and handed out by
            // one shuffle per edge and chunk, instead of one 4-byte load per lane, edge and chunk (measured 17.4 vs 9.3 ms per call for the
            // transposed aggregation of the Reddit-shaped graph against the single-head kernel).
            // Only through the transpose permutation: edge-ordered weights are sequential 32-byte runs that the per-lane loads already fetch
            // well (10.7 ms; the register form at three CTAs per SM measured 11.7), the permuted ones are random (17.4 -> 12.6 ms).
            constexpr bool H8 = MODE == M_EDGE_PERM_H && G >= 8;
            constexpr int EPR = G >= 8 ? G / 8 : 1;   // edges per register row
            const bool h8 = H8 && a.heads == 8;       // warp-uniform
            float wreg[H8 ? 8 : 1];
            if (h8) {
#pragma unroll
              for (int r = 0; r < 8; r++) {
                const int je = r * EPR + gl / 8;      // edge of the batch this lane fetches a weight of
                uint32_t eidx = b + (uint32_t)je;
                if (MODE == M_EDGE_PERM_H) eidx = __float_as_uint(__shfl_sync(gmask, w, je, G));
                wreg[r] = je < cnt ? __ldg(a.vals + (size_t)eidx * 8 + (gl & 7)) : 0.0f;
              }
            }
            if (cnt == G) {
#pragma unroll
              for (int j = 0; j < G; j += U) {
                float4 x[U][K];
                float wv[mode_has_heads<MODE>() ? U : 1][mode_has_heads<MODE>() ? K : 1];
#pragma unroll
                for (int u = 0; u < U; u++) {
                  const uint32_t cc = __shfl_sync(gmask, c_cur, j + u, G);
#pragma unroll
                  for (int k = 0; k < K; k++) x[u][k] = gather4(row_chunk<SPLIT>(bk[k], hk[k], a.n_split, row_bytes, cc));
                  if (mode_has_heads<MODE>() && !h8) {   // per-(edge, head) weights, requested with the gathers
                    const uint32_t eidx = MODE == M_EDGE_H ? b + (uint32_t)(j + u) : __float_as_uint(__shfl_sync(gmask, w, j + u, G));
#pragma unroll
                    for (int k = 0; k < K; k++) wv[u][k] = __ldg(a.vals + (size_t)eidx * a.heads + hd[k]);
                  }
                }
                if (H8 && h8) {
#pragma unroll
                  for (int u = 0; u < U; u++)
#pragma unroll
                    for (int k = 0; k < K; k++) wv[u][k] = __shfl_sync(gmask, wreg[(j + u) / EPR], ((j + u) % EPR) * 8 + hd[k], G);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                  // mean aggregation: every edge of the row carries the same 1/deg_i — no broadcast needed
                  const float ww = (MODE == M_MEAN || mode_has_heads<MODE>()) ? wrow : __shfl_sync(gmask, w, j + u, G);
#pragma unroll
                  for (int k = 0; k < K; k++) { float4 p; mul_w_f4(mode_has_heads<MODE>() ? wv[u][k] : ww, x[u][k], p); acc_add(acc[k], p); }
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < G; j += U) {
                if (j >= cnt) break;
                float4 x[U][K];
                float wv[mode_has_heads<MODE>() ? U : 1][mode_has_heads<MODE>() ? K : 1];
#pragma unroll
                for (int u = 0; u < U; u++) {
                  const uint32_t cc = __shfl_sync(gmask, c_cur, j + u, G);
#pragma unroll
                  for (int k = 0; k < K; k++)
                    x[u][k] = (j + u < cnt) ? gather4(row_chunk<SPLIT>(bk[k], hk[k], a.n_split, row_bytes, cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
                  if (mode_has_heads<MODE>() && !h8) {
                    const uint32_t eidx = MODE == M_EDGE_H ? b + (uint32_t)(j + u) : __float_as_uint(__shfl_sync(gmask, w, j + u, G));
#pragma unroll
                    for (int k = 0; k < K; k++) wv[u][k] = (j + u < cnt) ? __ldg(a.vals + (size_t)eidx * a.heads + hd[k]) : 0.0f;
                  }
                }
                if (H8 && h8) {
#pragma unroll
                  for (int u = 0; u < U; u++)
#pragma unroll
                    for (int k = 0; k < K; k++) wv[u][k] = __shfl_sync(gmask, wreg[(j + u) / EPR], ((j + u) % EPR) * 8 + hd[k], G);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                  const float ww = (MODE == M_MEAN || mode_has_heads<MODE>()) ? wrow : __shfl_sync(gmask, w, j + u, G);
                  if (j + u < cnt) {
#pragma unroll
                    for (int k = 0; k < K; k++) { float4 p; mul_w_f4(mode_has_heads<MODE>() ? wv[u][k] : ww, x[u][k], p); acc_add(acc[k], p); }
                  }
                }
              }
            }
            c_cur = c_nb;
          }
#pragma unroll
          for (int k = 0; k < K; k++)
            if (av[k]) store_chunk(a, row, cb + gl + G * k, acc[k]);
        }
      }
      c_first = c_nextrow;
    }
  }
}

// ---- hub rows: warp-specialised CTA, mbarrier ring ------------------------------------------------------------------
// 16 producer warps + 4 consumer warps. Producer warp w owns ring slot w and fills it for stages w, w+16, w+32, ...
// (stage k = edges [s + k*ES, s + (k+1)*ES) of the row): it loads the stage's column indices / weights with one
// coalesced request (prefetched one stage ahead), gathers the ES neighbour rows with up to 8 independent 128-bit loads in
// flight per lane, scales them and stores the PRODUCTS into its slot. The consumers (one thread per float4 column
// chunk) wait for the slots in stage order and add the products in edge order, so the fp32 result is the sequential
// sum the reference computes, while 16 stages are being gathered concurrently.
constexpr int HUB_CONS_WARPS = 4;
constexpr int HUB_CONS_THREADS = HUB_CONS_WARPS * 32;
constexpr int HUB_PROD_WARPS = 16;
constexpr int HUB_THREADS = HUB_CONS_THREADS + HUB_PROD_WARPS * 32;
constexpr int HUB_MAX_CHUNKS = 128;  // column block = 512 floats
#ifndef GAI_HUB_MIN_CTAS
#define GAI_HUB_MIN_CTAS 1
#endif
constexpr int HUB_MIN_CTAS = GAI_HUB_MIN_CTAS;
constexpr size_t HUB_RING_CAP = (HUB_MIN_CTAS == 1 ? 200 : 100) * 1024;

// ES = edges per stage (32, 16, 8 or 4). Dynamic smem: HUB_PROD_WARPS * ES * min(hub_per,128) float4.
// CL = lanes along the column-chunk dimension of one gather instruction (the other 32/CL lanes cover consecutive edges):
//   CL = 32  one edge per instruction, 32 chunks wide — column blocks of 32..128 chunks
//   CL = 8   four edges per instruction, 8 chunks (one 128-byte line per neighbour row) wide — the row's columns are split
//            into blocks of <= 8 chunks handled by gridDim.y CTAs: the sequential add chain of a 94 K-edge row stays
//            sequential per column, but its gather is spread over several SMs with every lane busy.
template <int MODE, int ES, int CL>
__global__ void __launch_bounds__(HUB_THREADS, HUB_MIN_CTAS) spmm_hub_kernel(const SpmmArgs a, const uint32_t* __restrict__ hub_rows) {
  extern __shared__ float4 ring[];
  __shared__ uint64_t full_bar[HUB_PROD_WARPS], empty_bar[HUB_PROD_WARPS];
  const uint32_t row = hub_rows[blockIdx.x];
  if (row < a.row_begin || row >= a.row_end) return;  // uniform for the CTA
  const uint32_t s = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
  const uint32_t nstages = (e - s + ES - 1) / ES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)a.ld_in * 4u;
  const char* inb = reinterpret_cast<const char*>(a.in);
  const char* halob = halo_virtual_base(a.in_halo, a.n_split, row_bytes);
  const int cb_begin = (int)blockIdx.y * a.hub_per;
  const int cb_end = a.nchunks < cb_begin + a.hub_per ? a.nchunks : cb_begin + a.hub_per;
  if (cb_begin >= cb_end) return;  // uniform for the CTA
  const int nch_max = (cb_end - cb_begin) < HUB_MAX_CHUNKS ? (cb_end - cb_begin) : HUB_MAX_CHUNKS;
  const size_t slot_stride = (size_t)ES * nch_max;
  uint32_t blk = 0;  // column-block index; slot p has been used blk * uses(p) times before this block (mbarrier phase bookkeeping)

  if (threadIdx.x == 0) {
    for (int i = 0; i < HUB_PROD_WARPS; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], HUB_CONS_WARPS); }
  }
  __syncthreads();

  for (int cb = cb_begin; cb < cb_end; cb += HUB_MAX_CHUNKS) {
    const int nch = (cb_end - cb) < HUB_MAX_CHUNKS ? (cb_end - cb) : HUB_MAX_CHUNKS;
    if (warp >= HUB_CONS_WARPS) {
      // ---------------- producers ----------------
      const int pw = warp - HUB_CONS_WARPS;
      const float wrow = (MODE == M_GCN || MODE == M_MEAN) ? __ldg(a.norm + row) : 0.0f;
      float4* slot = ring + (size_t)pw * slot_stride;
      const uint32_t round0 = blk * ((nstages + HUB_PROD_WARPS - 1 - pw) / HUB_PROD_WARPS);
      // prefetch the first stage's indices / weights (lane l < ES holds edge l of the stage)
      uint32_t c = 0; float w = 0.0f;
      {
        const uint32_t idx = s + (uint32_t)pw * ES + lane;
        if (lane < ES && idx < e) { c = __ldg(a.colidx + idx); w = edge_weight_t<MODE>(a, wrow, idx, c); }
      }
      for (uint32_t k = pw, r = 0; k < nstages; k += HUB_PROD_WARPS, r++) {
        const uint32_t base = s + k * ES;
        const int cnt = (e - base) < (uint32_t)ES ? (int)(e - base) : ES;
        const uint32_t cur_c = c; const float cur_w = w;
        // prefetch the next stage this warp owns
        c = 0; w = 0.0f;
        {
          const uint64_t nidx = (uint64_t)base + (uint64_t)HUB_PROD_WARPS * ES + lane;
          if (lane < ES && nidx < e) { c = __ldg(a.colidx + nidx); w = edge_weight_t<MODE>(a, wrow, (uint32_t)nidx, c); }
        }
        mbar_wait(&empty_bar[pw], ((round0 + r) & 1) ^ 1);
        constexpr int EL = 32 / CL;  // edges per gather instruction
        const int cl = lane % CL, el = lane / CL;
        for (int ch0 = 0; ch0 < nch; ch0 += CL) {
          const int ch = ch0 + cl;
          const bool chv = ch < nch;
          constexpr int UB = (ES / EL) < 8 ? (ES / EL) : 8;  // independent loads in flight per lane
#pragma unroll
          for (int j0 = 0; j0 < ES; j0 += UB * EL) {
            float4 x[UB]; float ww[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
              const int j = j0 + u * EL + el;
              const uint32_t cc = __shfl_sync(0xffffffffu, cur_c, j & 31);
              ww[u] = __shfl_sync(0xffffffffu, cur_w, j & 31);
              if (chv && j < cnt) x[u] = __ldg(row_chunk<true>(inb + (size_t)(cb + ch) * 16, halob + (size_t)(cb + ch) * 16, a.n_split, row_bytes, cc));
            }
#pragma unroll
            for (int u = 0; u < UB; u++) {
              const int j = j0 + u * EL + el;
              if (chv && j < cnt) {
                float4 p;
                p.x = __fmul_rn(ww[u], x[u].x); p.y = __fmul_rn(ww[u], x[u].y); p.z = __fmul_rn(ww[u], x[u].z); p.w = __fmul_rn(ww[u], x[u].w);
                slot[(size_t)j * nch + ch] = p;
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[pw]);
      }
    } else {
      // ---------------- consumers: in-order add, one float4 chunk per thread ----------------
      const int t = threadIdx.x;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (uint32_t k = 0; k < nstages; k++) {
        const int pw = k % HUB_PROD_WARPS;
        const uint32_t r = k / HUB_PROD_WARPS;
        const uint32_t round0 = blk * ((nstages + HUB_PROD_WARPS - 1 - pw) / HUB_PROD_WARPS);
        const uint32_t base = s + k * ES;
        const int cnt = (e - base) < (uint32_t)ES ? (int)(e - base) : ES;
        mbar_wait(&full_bar[pw], (round0 + r) & 1);
        if (t < nch) {
          const float4* tile = ring + (size_t)pw * slot_stride + t;
          if (cnt == ES) {
            // software pipeline in sub-blocks of SB edges: the shared-memory loads of sub-block i+1 are issued before the
            // dependent add chain of sub-block i, so only the first load latency of a stage is exposed
            constexpr int SB = ES < 8 ? ES : 8;
            float4 p[2][SB];
#pragma unroll
            for (int j = 0; j < SB; j++) p[0][j] = tile[(size_t)j * nch];
#pragma unroll
            for (int b = 0; b < ES / SB; b++) {
              if (b + 1 < ES / SB) {
#pragma unroll
                for (int j = 0; j < SB; j++) p[(b + 1) & 1][j] = tile[(size_t)((b + 1) * SB + j) * nch];
              }
#pragma unroll
              for (int j = 0; j < SB; j++) {
                const float4 q = p[b & 1][j];
                acc.x = __fadd_rn(acc.x, q.x); acc.y = __fadd_rn(acc.y, q.y); acc.z = __fadd_rn(acc.z, q.z); acc.w = __fadd_rn(acc.w, q.w);
              }
            }
          } else {
            for (int j = 0; j < cnt; j++) {
              const float4 p = tile[(size_t)j * nch];
              acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y); acc.z = __fadd_rn(acc.z, p.z); acc.w = __fadd_rn(acc.w, p.w);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[pw]);
      }
      if (t < nch) store_chunk(a, row, cb + t, acc);
    }
    blk++;
  }
}

// in [n x F] (ld_in) -> padded [n x 4*nchunks], zero-filled tail columns
__global__ void pad_rows_kernel(size_t n_rows, int F, int Fp, const float* __restrict__ in, int ld_in, float* __restrict__ out) {
  const size_t total = n_rows * (size_t)Fp;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / Fp;
    const int c = (int)(i % Fp);
    out[i] = c < F ? __ldg(in + r * ld_in + c) : 0.0f;
  }
}

inline bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % 16) == 0; }

// Work list serving a call: the whole-graph list, the list of a registered segment (gai_csr_set_row_segments) whose bounds
// equal the requested range, or none (natural order, 32 consecutive rows per claim).
struct ListSel { const uint32_t* order; const uint32_t* claim_ptr; const uint32_t* hub_rows; uint32_t n_hub; unsigned long long n_claims; bool listed; };
ListSel select_list(const SpmmArgs& a, const gai_csr* g) {
  if (a.row_begin == 0 && a.row_end == g->nv && g->row_order != nullptr)
    return {g->row_order + g->n_hub, g->claim_ptr, g->hub_rows, g->n_hub, g->n_claims, true};
  for (int i = 0; i < g->n_seg; i++)
    if (g->seg[i].rb == a.row_begin && g->seg[i].re == a.row_end && g->seg[i].row_order != nullptr)
      return {g->seg[i].row_order + g->seg[i].n_hub, g->seg[i].claim_ptr, g->seg[i].row_order, g->seg[i].n_hub, g->seg[i].n_claims, true};
  const unsigned long long n_rows = (unsigned long long)a.row_end - a.row_begin;
  return {nullptr, nullptr, g->hub_rows, g->n_hub, (n_rows + SLOTS - 1) / SLOTS, false};
}

// float4 chunks per hub column block: 4 once the longest row passes 128 K edges (its add chain then bounds the kernel and edges in flight
// per ring byte are what shortens it: N = 8 shard of configs[1], F = 47: 3.27 -> 2.84 ms), else 8. GAI_HUB_CL=4|8 overrides (experiments).
inline int hub_block_chunks(const gai_csr* g) {
  static const int forced = [] { const char* e = getenv("GAI_HUB_CL"); const int v = e ? atoi(e) : 0; return (v == 4 || v == 8) ? v : 0; }();
  if (forced) return forced;
  return g->max_degree > 131072u ? 4 : 8;
}
// widths whose hub rows are handled inside the persistent kernel (column blocks of <= 4 chunks, up to 32 blocks)
inline bool hub_fused(const SpmmArgs& a) { return a.nchunks <= 128; }

template <int MODE>
int launch_rows_mode(SpmmArgs a, const gai_csr* g, cudaStream_t st) {
  // listed calls walk a degree-ordered list (hub rows sit at its head and are skipped by offset); other row-range calls walk
  // the range in natural order
  const ListSel sel = select_list(a, g);
  const uint32_t* order = sel.order;
  const unsigned long long claims = sel.n_claims;
  unsigned long long hub_items = 0;
  int nsplit = 1;
  if (sel.n_hub != 0 && hub_fused(a)) {
    a.hub_cl = hub_block_chunks(g);
    nsplit = (a.nchunks + a.hub_cl - 1) / a.hub_cl;
    a.hub_per = (a.nchunks + nsplit - 1) / nsplit;
    hub_items = (unsigned long long)sel.n_hub * (unsigned)nsplit;
  }
  // GAI_SPMM_DEBUG (timing experiments only, results are wrong): 1 = skip the hub items, 2 = skip the light rows — what each half of the
  // persistent kernel costs on a given graph (tools/shard_probe.py)
  static const int debug_knob = [] {
    const int k = getenv("GAI_SPMM_DEBUG") ? atoi(getenv("GAI_SPMM_DEBUG")) : 0;
    if (k) fprintf(stderr, "libgai_b200: GAI_SPMM_DEBUG=%d — timing experiment, aggregation RESULTS ARE WRONG\n", k);
    return k;
  }();
  unsigned long long claims_run = claims;
  if (debug_knob == 1) hub_items = 0;
  if (debug_knob == 2) claims_run = 0;
  if (claims_run == 0 && hub_items == 0) return GAI_OK;
  const uint32_t* claim_ptr = sel.claim_ptr;
  int G = 4;
  while (G < 32 && G < a.nchunks) G <<= 1;
  int K = 1;
  if (G == 32) { K = (a.nchunks + 31) / 32; K = K <= 1 ? 1 : (K <= 2 ? 2 : 4); }
  unsigned long long ctas = (claims + 7) / 8;
  if (ctas < hub_items) ctas = hub_items;
  const bool share = (a.flags & GAI_SPMM_SHARE_SMS) != 0;   // leave registers / threads for one foreign CTA per SM
  const int resident = MODE == M_EDGE_PERM_H ? 3 : 4;   // CTAs per SM (launch bounds of the kernel)
  const unsigned long long persistent = (unsigned long long)gai::sm_count() * (share ? resident - 1 : resident);
  if (ctas > persistent) ctas = persistent;
  const unsigned grid = (unsigned)ctas;
  // rotating work-counter pairs {light-row claims, hub items}: launches on one stream are ordered; the rotation keeps up to 8
  // launches that overlap on different streams (interior / boundary rows of the 1D partition) from sharing a pair
  unsigned long long* ctr = g->row_counters + 2 * (__atomic_fetch_add(&const_cast<gai_csr*>(g)->counter_seq, 1u, __ATOMIC_RELAXED) % 8u);
  GAI_CUDA(cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), st));
#define GAI_ROWS_LAUNCH(GG, KK)                                                                                                        \
  do {                                                                                                                                \
    if (a.in_halo) spmm_rows_kernel<MODE, GG, KK, true><<<grid, 256, 0, st>>>(a, order, claim_ptr, claims_run, ctr, sel.hub_rows, hub_items, nsplit); \
    else spmm_rows_kernel<MODE, GG, KK, false><<<grid, 256, 0, st>>>(a, order, claim_ptr, claims_run, ctr, sel.hub_rows, hub_items, nsplit);          \
  } while (0)
  if (G == 4) GAI_ROWS_LAUNCH(4, 1);
  else if (G == 8) GAI_ROWS_LAUNCH(8, 1);
  else if (G == 16) GAI_ROWS_LAUNCH(16, 1);
  else if (K == 1) GAI_ROWS_LAUNCH(32, 1);
  else if (K == 2) GAI_ROWS_LAUNCH(32, 2);
  else GAI_ROWS_LAUNCH(32, 4);
#undef GAI_ROWS_LAUNCH
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

int launch_rows(const SpmmArgs& a, const gai_csr* g, cudaStream_t st) {
  switch (a.mode) {
    case M_GCN: return launch_rows_mode<M_GCN>(a, g, st);
    case M_MEAN: return launch_rows_mode<M_MEAN>(a, g, st);
    case M_MEAN_T: return launch_rows_mode<M_MEAN_T>(a, g, st);
    case M_EDGE: return launch_rows_mode<M_EDGE>(a, g, st);
    case M_EDGE_H: return launch_rows_mode<M_EDGE_H>(a, g, st);
    case M_EDGE_PERM_H: return launch_rows_mode<M_EDGE_PERM_H>(a, g, st);
    default: return launch_rows_mode<M_EDGE_PERM>(a, g, st);
  }
}

template <int MODE, int ES, int CL>
int launch_hub_es(const SpmmArgs& a, const gai_csr* g, size_t smem, unsigned nsplit, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    GAI_CUDA(cudaFuncSetAttribute(spmm_hub_kernel<MODE, ES, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    configured = true;
  }
  const ListSel sel = select_list(a, g);
  spmm_hub_kernel<MODE, ES, CL><<<dim3(sel.n_hub, nsplit), HUB_THREADS, smem, st>>>(a, sel.hub_rows);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

template <int MODE>
int launch_hub_mode(SpmmArgs a, const gai_csr* g, cudaStream_t st) {
  // only widths beyond 512 floats come here (narrower rows are hub items of the persistent kernel)
  a.hub_per = a.nchunks;
  const int nch = a.nchunks < HUB_MAX_CHUNKS ? a.nchunks : HUB_MAX_CHUNKS;
  // largest stage size in {32, 16, 8, 4} edges whose 16-slot ring fits 200 KB of shared memory (longer stages amortise the
  // consumer's per-stage barrier round trip over more in-order adds)
  auto bytes = [&](int es) { return (size_t)HUB_PROD_WARPS * es * nch * sizeof(float4); };
  if (bytes(32) <= HUB_RING_CAP) return launch_hub_es<MODE, 32, 32>(a, g, bytes(32), 1, st);
  if (bytes(16) <= HUB_RING_CAP) return launch_hub_es<MODE, 16, 32>(a, g, bytes(16), 1, st);
  if (bytes(8) <= HUB_RING_CAP) return launch_hub_es<MODE, 8, 32>(a, g, bytes(8), 1, st);
  return launch_hub_es<MODE, 4, 32>(a, g, bytes(4), 1, st);
}

int launch_hub(const SpmmArgs& a, const gai_csr* g, cudaStream_t st) {
  if (select_list(a, g).n_hub == 0) return GAI_OK;
  switch (a.mode) {
    case M_GCN: return launch_hub_mode<M_GCN>(a, g, st);
    case M_MEAN: return launch_hub_mode<M_MEAN>(a, g, st);
    case M_MEAN_T: return launch_hub_mode<M_MEAN_T>(a, g, st);
    case M_EDGE: return launch_hub_mode<M_EDGE>(a, g, st);
    default: return launch_hub_mode<M_EDGE_PERM>(a, g, st);
  }
}

int spmm_dispatch(gai_csr_t g, int mode, uint32_t rb, uint32_t re, int F, const float* vals, const uint32_t* perm, const float* in,
                  int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream, const uint32_t* mask_bits = nullptr,
                  int ld_bits = 0, const float* in_halo = nullptr, uint32_t n_split = 0xffffffffu, int heads = 1) {
  GAI_CHECK_ARG(g != nullptr);
  GAI_CHECK_ARG(rb <= re && re <= g->nv);
  if (re == rb) return GAI_OK;  // empty graph / empty row range: nothing to do (buffers may be NULL)
  GAI_CHECK_ARG(in != nullptr && out != nullptr);
  GAI_CHECK_ARG(F > 0 && ld_in >= F && ld_out >= F && ld_in < (1 << 28));  // row pitch in bytes is 32-bit inside the kernels
  GAI_CHECK_ARG(!(flags & GAI_EPI_ADD) || addend != nullptr);
  GAI_CHECK_ARG(mode < M_EDGE || vals != nullptr);
  if (heads > 1) {
    // multi-head edge values: a power-of-two number of heads, whole float4 chunks per head, a power-of-two number of them
    const int cols = F / heads;
    GAI_CHECK_ARG((mode == M_EDGE || mode == M_EDGE_PERM) && F % heads == 0 && cols % 4 == 0 && ((cols / 4) & (cols / 4 - 1)) == 0 && F <= 512);
    mode = mode == M_EDGE ? M_EDGE_H : M_EDGE_PERM_H;
  }
  GAI_CHECK_ARG(in != out);
  cudaStream_t st = gai::S(stream);
  SpmmArgs a;
  a.rowptr = g->rowptr; a.colidx = g->colidx;
  a.norm = (mode == M_GCN) ? g->norm_gcn : g->norm_mean;
  a.vals = vals; a.perm = perm; a.out = out; a.addend = addend;
  a.F = F; a.nchunks = (F + 3) / 4; a.ld_out = ld_out; a.row_begin = rb; a.row_end = re;
  a.mode = mode; a.flags = flags;
  a.hub_threshold = g->n_hub ? g->hub_degree : 0xffffffffu;
  a.hub_per = a.nchunks; a.hub_cl = 8;
  a.mask_bits = mask_bits; a.ld_bits = ld_bits;
  a.in_halo = in_halo; a.n_split = in_halo ? n_split : 0xffffffffu;
  a.heads = heads; a.hshift = 0;
  if (heads > 1) { int cph = F / heads / 4; while (cph > 1) { a.hshift++; cph >>= 1; } }
  a.out_vec = (F % 4 == 0) && (ld_out % 4 == 0) && aligned16(out) && aligned16(addend);
  if (in_halo && !((ld_in % 4 == 0) && aligned16(in) && aligned16(in_halo) && ld_in >= a.nchunks * 4))
    return gai::set_error(GAI_ERR_ARG, "spmm", "a split (masters | halo) input needs 16-byte aligned rows whose pitch is a multiple of 4 floats");
  if ((ld_in % 4 == 0) && aligned16(in) && ld_in >= a.nchunks * 4) {
    // rows are 128-bit loadable as stored; when F % 4 != 0 the tail chunk also reads the (ld_in - F) padding columns of
    // the row: they land in accumulator lanes that are never stored (store_chunk masks columns >= F)
    a.in = in; a.ld_in = ld_in;
  } else {
    // gather source must be 128-bit loadable: stage a zero-padded copy (all nv rows can be neighbours)
    const int Fp = a.nchunks * 4;
    void* ws = nullptr;
    int rc = gai::workspace_slot(1, sizeof(float) * (size_t)g->nv * Fp, &ws, st);
    if (rc != GAI_OK) return rc;
    const size_t total = (size_t)g->nv * Fp;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)gai::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    pad_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(g->nv, F, Fp, in, ld_in, reinterpret_cast<float*>(ws));
    GAI_LAUNCH_CHECK();
    a.in = reinterpret_cast<const float*>(ws); a.ld_in = Fp;
  }
  // Widths beyond 512 floats only: the hub rows of such calls run in the CTA-per-row kernel on a high-priority side
  // stream (fork/join with events); every narrower call serves its hub rows as work items of the persistent kernel.
  int rc = GAI_OK;
  const bool has_hub = select_list(a, g).n_hub != 0 && !hub_fused(a);
  if (has_hub) {
    if (!g->aux_stream) {
      int lo = 0, hi = 0;
      GAI_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      GAI_CUDA(cudaStreamCreateWithPriority(&g->aux_stream, cudaStreamNonBlocking, hi));
      GAI_CUDA(cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming));
      GAI_CUDA(cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming));
    }
    GAI_CUDA(cudaEventRecord(g->ev_fork, st));
    GAI_CUDA(cudaStreamWaitEvent(g->aux_stream, g->ev_fork, 0));
    rc = launch_hub(a, g, g->aux_stream);
    if (rc != GAI_OK) return rc;
    GAI_CUDA(cudaEventRecord(g->ev_join, g->aux_stream));
  }
  rc = launch_rows(a, g, st);
  if (has_hub) GAI_CUDA(cudaStreamWaitEvent(st, g->ev_join, 0));
  return rc;
}

}  // namespace

extern "C" {

int gai_spmm_gcn(gai_csr_t g, int F, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr);
  return spmm_dispatch(g, M_GCN, 0, g->nv, F, nullptr, nullptr, in, ld_in, out, ld_out, flags, addend, stream);
}
int gai_spmm_mean(gai_csr_t g, int F, const float* in, int ld_in, float* out, int ld_out, int transposed, int flags, const float* addend, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr);
  return spmm_dispatch(g, transposed ? M_MEAN_T : M_MEAN, 0, g->nv, F, nullptr, nullptr, in, ld_in, out, ld_out, flags, addend, stream);
}
int gai_spmm_edge(gai_csr_t g, int F, const float* vals, const uint32_t* perm, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr);
  return spmm_dispatch(g, perm ? M_EDGE_PERM : M_EDGE, 0, g->nv, F, vals, perm, in, ld_in, out, ld_out, flags, addend, stream);
}
int gai_spmm_edge_heads(gai_csr_t g, int F, int heads, const float* vals, const uint32_t* perm, const float* in, int ld_in, float* out, int ld_out, int flags,
                        const float* addend, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr && heads >= 1);
  return spmm_dispatch(g, perm ? M_EDGE_PERM : M_EDGE, 0, g->nv, F, vals, perm, in, ld_in, out, ld_out, flags, addend, stream, nullptr, 0, nullptr,
                       0xffffffffu, heads);
}
int gai_spmm_gcn_masked(gai_csr_t g, int F, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend,
                        const uint32_t* mask_bits, int ld_bits, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr && mask_bits != nullptr && ld_bits >= (F + 31) / 32);
  return spmm_dispatch(g, M_GCN, 0, g->nv, F, nullptr, nullptr, in, ld_in, out, ld_out, flags, addend, stream, mask_bits, ld_bits);
}
int gai_spmm_mean_masked(gai_csr_t g, int F, const float* in, int ld_in, float* out, int ld_out, int transposed, int flags, const float* addend,
                         const uint32_t* mask_bits, int ld_bits, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr && mask_bits != nullptr && ld_bits >= (F + 31) / 32);
  return spmm_dispatch(g, transposed ? M_MEAN_T : M_MEAN, 0, g->nv, F, nullptr, nullptr, in, ld_in, out, ld_out, flags, addend, stream, mask_bits,
                       ld_bits);
}
int gai_spmm_rows_ex(gai_csr_t g, int mode, uint32_t rb, uint32_t re, int F, const float* vals, const uint32_t* perm, const float* in, int ld_in,
                     float* out, int ld_out, int flags, const float* addend, const uint32_t* mask_bits, int ld_bits, const float* in_halo,
                     uint32_t n_split, gai_stream_t stream) {
  GAI_CHECK_ARG(mode >= M_GCN && mode <= M_EDGE_PERM);
  GAI_CHECK_ARG(mode != M_EDGE_PERM || perm != nullptr);
  GAI_CHECK_ARG(mask_bits == nullptr || ld_bits >= (F + 31) / 32);
  return spmm_dispatch(g, mode, rb, re, F, vals, perm, in, ld_in, out, ld_out, flags, addend, stream, mask_bits, ld_bits, in_halo, n_split);
}
int gai_spmm_gcn_rows(gai_csr_t g, uint32_t rb, uint32_t re, int F, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream) {
  return spmm_dispatch(g, M_GCN, rb, re, F, nullptr, nullptr, in, ld_in, out, ld_out, flags, addend, stream);
}
int gai_spmm_mean_rows(gai_csr_t g, uint32_t rb, uint32_t re, int F, const float* in, int ld_in, float* out, int ld_out, int transposed, int flags, const float* addend, gai_stream_t stream) {
  return spmm_dispatch(g, transposed ? M_MEAN_T : M_MEAN, rb, re, F, nullptr, nullptr, in, ld_in, out, ld_out, flags, addend, stream);
}

}  // extern "C"
