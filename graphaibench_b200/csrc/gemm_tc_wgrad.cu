// Weight gradient on the 5th-generation tensor cores: C[Kx x My] = A^T · B with A [n x Kx] and B [n x My] row-major and
// the reduction running over the n (vertex) rows — dW = X^T·G of every layer (src/gnn/gconv/gcn_layer.cpp:50,56,
// sage_layer.cpp:37-47, gat_layer.cpp:38; matmul(..., transA = true), src/utilities/math_functions.cpp:142-171).
//
// Both operands are "MN-major" for the MMA (the reduction index is the slow axis in memory), which tcgen05 kind::tf32
// takes directly from 128-byte-swizzled shared memory, so the activations are streamed exactly once, in their stored layout:
//   * TMA box = [16 rows x 32 columns] (one 128-byte swizzle span per row, 2 KB); a k-block is 16 rows of A (4 boxes per
//     128-column M tile) and of B (ceil(My/32) boxes). Columns past the matrix edge are zero-filled by TMA.
//   * 3xTF32: splitter warps rewrite each landed box as hi = rn_tf32(x) in place and lo = rn_tf32(x - hi) beside it;
//     the MMA warp issues A_lo·B_hi + A_hi·B_lo + A_hi·B_hi (fp32 accumulation in TMEM) for each 8-row k-group.
//   * split over the rows: a persistent grid of one CTA per SM, each CTA reduces a contiguous range of rows into its own
//     TMEM accumulators (1 or 2 M tiles x up to 256 columns = up to all 512 TMEM columns), writes one fp32 partial, and a
//     second kernel adds the partials in CTA order (deterministic, no atomics).
//   * concatenated forms (WgradCat), one pass over the shared operand instead of two:
//       two A, shared B   [dWn; dWs] = [ÂX | X]^T · dH   each A part fills one 128-column M tile (own TMA map)
//       shared A, two B   [dWn | dWs] = H^T · [dY | dZ]   the B parts sit side by side in the N dimension (own TMA maps)
//     (SAGE keeps W_neigh and W_self apart, sage_layer.cpp:37-47; the reference runs two sgemm calls.)
// HBM-bound by design: bytes = 4·n·(Kx + My), tensor work = 3 · 2·n·128·ceil(Kx/128)·My.
#include <cstdlib>
#include "tc_common.cuh"

namespace gai {

namespace {

using namespace tc;

constexpr int WG_BK = 16;                 // rows per k-block (8-row blocks measured 40 % slower: twice the TMA boxes and barrier hand-offs per byte)
constexpr uint32_t WG_BOX = WG_BK * 128;  // bytes per TMA box
constexpr uint32_t WG_SEG_KB = 2048 / WG_BK;  // k-blocks (2048 rows) one TMEM accumulator absorbs before it is flushed into the CTA's fp32 partial: the
                                          // tensor core adds into TMEM with truncation, so the error of one accumulator grows linearly with the k-steps
                                          // it absorbs. One accumulator per CTA over 2.45 M rows (16.5 K rows each) measured 0.9-1.3e-4 of the result
                                          // norm against the exact fp64 product — 20x the reference's OpenBLAS sgemm; segments of 2048 rows: 1-2e-5
                                          // (tests/test_reference_parity_gpu.py). Each flush adds the segment into the CTA's L2-resident partial with
                                          // ordinary round-to-nearest fp32 adds, the old values of a 32-column chunk requested together before the
                                          // TMEM read-back (one dependent load per value measured 35 us per flush; one partial slot per segment
                                          // cost as much in memset + write + re-read traffic: 3 x 242 MB per call). The reduce kernel adds the
                                          // CTAs' partials in double.
constexpr int WG_THREADS = 384;           // warp 0 TMA, warp 1 MMA, warps 4-11 splitters + epilogue
constexpr int WG_SPLIT_WARPS = 8;

struct WgArgs {
  int debug;  // GAI_TC_DEBUG (timing experiments only): 4 = no hi/lo split, 8 = no MMA issue
  float* partial;  // [grid][Kx_total][My_total], parts concatenated
  size_t nrows;
  size_t blocks_per_cta;  // k-blocks per CTA
  int Kx, My;      // totals over the parts
  int mt;          // 128-column tiles of A (1 or 2)
  int n_mma;       // tile width (multiple of 32)
  int dual_a;      // 1: M tile t streams from map_a[t]; 0: both tiles are column ranges of map_a[0]
  int nbox_at[2];  // live 32-column boxes of each M tile
  int rows_t[2];   // live accumulator rows of each M tile
  int nbox_b, nbox_b0;  // B boxes; the first nbox_b0 stream from map_b[0], the rest from map_b[1]
  int my0;         // live columns of the first B part (the second starts at tile column 32*nbox_b0)
  int stages, passes;
  int ldp;         // row pitch of the partial: each B part's columns padded to 4 floats (part 1 starts at my0p)
  int my0p;        // my0 rounded up to 4: a lane's 32-column run of either part is then 16-byte aligned
  uint32_t a_bytes, b_bytes, stage_bytes;
};

// MN-major tf32 operand = layout type SWIZZLE_128B_BASE32B (the only one tcgen05 takes for 32-bit MN-major data): atoms
// of [32 fp32 along M/N] x [4 rows along K] (512 B; 32-byte chunks XOR-swizzled by row, what TMA's SWIZZLE_128B_ATOM_32B
// writes), the next 32 columns one box further (leading byte offset), the next 4 rows 512 B further (stride byte offset).
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(WG_BOX >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)1 << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                     const __grid_constant__ CUtensorMap map_b0, const __grid_constant__ CUtensorMap map_b1, const WgArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], conv_bar[8], empty_bar[8], done_bar, tfree_bar;
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t total_kb = (g.nrows + WG_BK - 1) / WG_BK;
  const size_t kb0 = (size_t)blockIdx.x * g.blocks_per_cta;
  const size_t kb1 = kb0 + g.blocks_per_cta < total_kb ? kb0 + g.blocks_per_cta : total_kb;
  const uint32_t nkb = (uint32_t)(kb1 - kb0);  // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; i++) { mbar_init(&full_bar[i], 1); mbar_init(&conv_bar[i], WG_SPLIT_WARPS); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    mbar_init(&tfree_bar, (uint32_t)g.mt * 4u);  // the warps that flush the accumulator
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ---------------- TMA producer: lane 0 owns the barriers, every box of a stage is issued by its own lane ----------------
    {
      const int nA0 = g.nbox_at[0], nA1 = g.nbox_at[1], nB = g.nbox_b;
      for (uint32_t it = 0; it < nkb; it++) {
        const int s = it % g.stages;
        if (lane == 0) {
          mbar_wait(&empty_bar[s], ((it / g.stages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(nA0 + nA1 + nB) * WG_BOX);
        }
        __syncwarp();
        uint8_t* st = smem + (size_t)s * g.stage_bytes;
        uint8_t* sb = st + 2 * g.a_bytes;
        const int row = (int)((kb0 + it) * WG_BK);
        if (lane < nA0) {
          tma_load_2d(st + (size_t)lane * WG_BOX, &map_a0, lane * 32, row, &full_bar[s]);
        } else if (lane < nA0 + nA1) {
          const int c = lane - nA0;
          tma_load_2d(st + (size_t)(4 + c) * WG_BOX, g.dual_a ? &map_a1 : &map_a0, (g.dual_a ? c : 4 + c) * 32, row, &full_bar[s]);
        } else if (lane < nA0 + nA1 + nB) {
          const int c = lane - nA0 - nA1;
          if (c < g.nbox_b0) tma_load_2d(sb + (size_t)c * WG_BOX, &map_b0, c * 32, row, &full_bar[s]);
          else tma_load_2d(sb + (size_t)c * WG_BOX, &map_b1, (c - g.nbox_b0) * 32, row, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      // D = F32, A = B = TF32, both MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(g.n_mma >> 3) << 17) | ((128u >> 4) << 24);
      for (uint32_t it = 0; it < nkb; it++) {
        const int s = it % g.stages;
        const uint32_t seg = it / WG_SEG_KB;
        const bool seg_start = it % WG_SEG_KB == 0;
        if (seg_start && seg > 0) mbar_wait(&tfree_bar, (seg - 1) & 1);  // the previous segment's sums have left the accumulator
        mbar_wait(&conv_bar[s], (it / g.stages) & 1);
        tcgen05_fence_after();
        const uint32_t a_hi = smem_u32(smem + (size_t)s * g.stage_bytes);
        const uint32_t a_lo = a_hi + g.a_bytes;
        const uint32_t b_hi = a_hi + 2 * g.a_bytes;
        const uint32_t b_lo = b_hi + g.b_bytes;
#pragma unroll
        for (int kg = 0; kg < ((g.debug & 8) ? 0 : WG_BK / 8); kg++) {
          const uint32_t koff = kg * 1024;  // 8 rows x 128 B
          for (int t = 0; t < g.mt; t++) {
            const uint32_t toff = (uint32_t)t * 4u * WG_BOX + koff;
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            const uint32_t first = (seg_start && kg == 0) ? 0u : 1u;
            if (g.passes == 3) {
              umma_tf32(d_tmem, make_desc_mn128(a_lo + toff), make_desc_mn128(b_hi + koff), idesc, first);
              umma_tf32(d_tmem, make_desc_mn128(a_hi + toff), make_desc_mn128(b_lo + koff), idesc, 1u);
              umma_tf32(d_tmem, make_desc_mn128(a_hi + toff), make_desc_mn128(b_hi + koff), idesc, 1u);
            } else {
              umma_tf32(d_tmem, make_desc_mn128(a_hi + toff), make_desc_mn128(b_hi + koff), idesc, first);
            }
          }
        }
        umma_commit(&empty_bar[s]);
        if ((it + 1) % WG_SEG_KB == 0 || it + 1 == nkb) umma_commit(&done_bar);  // segment complete
      }
    }
  } else if (warp >= 4) {
    // ---------------- splitters: x -> (rn_tf32(x) in place, rn_tf32(x - hi)) for the A and B boxes ----------------
    const int t = threadIdx.x - 128;  // 0..255
    const uint32_t a_u4 = g.a_bytes / 16, b_u4 = g.b_bytes / 16;
    // zero the A boxes TMA never fills (the tail boxes of each M tile), once per stage buffer
    constexpr uint32_t BOX_U4 = WG_BOX / 16;
    if (g.nbox_at[0] < 4 || (g.mt > 1 && g.nbox_at[1] < 4)) {
      for (int s = 0; s < g.stages; s++) {
        uint4* base = reinterpret_cast<uint4*>(smem + (size_t)s * g.stage_bytes);
        for (int tl = 0; tl < g.mt; tl++) {
          const uint32_t z0 = (uint32_t)(tl * 4 + g.nbox_at[tl]) * BOX_U4, z1 = (uint32_t)(tl * 4 + 4) * BOX_U4;
          for (uint32_t i = z0 + t; i < z1; i += WG_SPLIT_WARPS * 32) { base[i] = make_uint4(0, 0, 0, 0); base[a_u4 + i] = make_uint4(0, 0, 0, 0); }
        }
      }
      fence_proxy_async();
    }
    for (uint32_t it = 0; it < nkb; it++) {
      const int s = it % g.stages;
      mbar_wait(&full_bar[s], (it / g.stages) & 1);
      if (g.passes == 3 && !(g.debug & 4)) {
        uint4* ahi = reinterpret_cast<uint4*>(smem + (size_t)s * g.stage_bytes);
        uint4* bhi = ahi + 2 * a_u4;
        const uint32_t a_live0 = (uint32_t)g.nbox_at[0] * BOX_U4;
        const uint32_t a_live = a_live0 + (uint32_t)g.nbox_at[1] * BOX_U4;
        for (uint32_t i = t; i < a_live + b_u4; i += WG_SPLIT_WARPS * 32) {
          // live A boxes: [0, nbox_at[0]) of tile 0, then [4, 4 + nbox_at[1]) of tile 1
          uint4* hi = i < a_live0 ? ahi + i : (i < a_live ? ahi + (i - a_live0) + 4 * BOX_U4 : bhi + (i - a_live));
          uint4* lo = i < a_live ? hi + a_u4 : hi + b_u4;
          const uint4 v = *hi;
          uint4 h, l;
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
          *hi = h;
          *lo = l;
        }
        fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv_bar[s]);
      // ---------------- flush: TMEM accumulator -> the partial slot of this (CTA, segment) ----------------
      if ((it + 1) % WG_SEG_KB == 0 || it + 1 == nkb) {
        const uint32_t seg = it / WG_SEG_KB;
        const int tile = (warp - 4) >> 2;  // warps 4-7: M tile 0, warps 8-11: M tile 1
        const int q = warp & 3;            // TMEM lane quarter this warp may read
        if (tile < g.mt) {
          mbar_wait(&done_bar, seg & 1);
          tcgen05_fence_after();
          const int kl = q * 32 + lane;  // accumulator row within the tile
          const bool live = kl < g.rows_t[tile];
          const int kx = (tile ? g.rows_t[0] : 0) + kl;  // row of the concatenated partial
          float* prow = g.partial + ((size_t)blockIdx.x * g.Kx + (live ? kx : 0)) * g.ldp;
          for (int c0 = 0; c0 < g.n_mma; c0 += 32) {
            // tile columns -> concatenated partial columns: part 0 at [0, my0), part 1 from tile column 32*nbox_b0
            const int part1 = c0 >= 32 * g.nbox_b0;
            const int pc0 = part1 ? g.my0p + c0 - 32 * g.nbox_b0 : c0;
            const int lim = part1 ? g.my0p + (g.My - g.my0) : g.my0;
            const bool vec = true;  // pc0 is a multiple of 4 by construction (my0p, 32-column chunks) and rows are pitched to 4 floats
            float4 old[8];
            if (seg && live && vec) {  // the running sums of this chunk, all eight requests in flight together
#pragma unroll
              for (int j = 0; j < 8; j++) old[j] = (pc0 + 4 * j + 3 < lim) ? *reinterpret_cast<const float4*>(prow + pc0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)tile * 256u + (uint32_t)c0, r);
            if (!live) continue;
            if (vec) {
#pragma unroll
              for (int j = 0; j < 8; j++) {
                if (pc0 + 4 * j + 3 < lim) {
                  float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                  if (seg) { v.x = __fadd_rn(v.x, old[j].x); v.y = __fadd_rn(v.y, old[j].y); v.z = __fadd_rn(v.z, old[j].z); v.w = __fadd_rn(v.w, old[j].w); }
                  *reinterpret_cast<float4*>(prow + pc0 + 4 * j) = v;
                } else {
#pragma unroll
                  for (int k = 0; k < 4; k++)
                    if (pc0 + 4 * j + k < lim) prow[pc0 + 4 * j + k] = seg ? __fadd_rn(prow[pc0 + 4 * j + k], __uint_as_float(r[4 * j + k])) : __uint_as_float(r[4 * j + k]);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (pc0 + j < lim) prow[pc0 + j] = seg ? __fadd_rn(prow[pc0 + j], __uint_as_float(r[j])) : __uint_as_float(r[j]);
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tfree_bar);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// C = (accum ? C : 0) + sum_p partial[p], p ascending (deterministic). The concatenated partial [Kx x My] is cut back into
// its destinations: rows >= kx0 belong to C1 (two-A form), columns >= my0 belong to C1 (two-B form).
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C0, size_t ldc0, float* __restrict__ C1, size_t ldc1,
                                    int Kx, int My, int ldp, int kx0, int my0, int parts, int accum) {
  // 8 lanes per output element, each adding every 8th partial in double; a butterfly over the 8 lanes finishes the sum (fixed order:
  // the same bits run to run).
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 3, slice = t & 7;
  const bool live = i < Kx * My;
  const int m = live ? i / My : 0, n = live ? i % My : 0;
  const int my0p = (my0 + 3) / 4 * 4;
  const size_t stride = (size_t)Kx * ldp, off = (size_t)m * ldp + (n >= my0 ? my0p + (n - my0) : n);
  double r = 0.0;
  if (live)
    for (int p = slice; p < parts; p += 8) r += (double)partial[(size_t)p * stride + off];
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  if (!live || slice != 0) return;
  float* dst;
  if (m >= kx0) dst = C1 + (size_t)(m - kx0) * ldc1 + n;
  else if (n >= my0) dst = C1 + (size_t)m * ldc1 + (n - my0);
  else dst = C0 + (size_t)m * ldc0 + n;
  *dst = (float)(r + (accum ? (double)*dst : 0.0));
}

// [n x F] (ld) -> [n x Fp], zero-filled tail columns (operands whose row pitch is not a multiple of 16 bytes)
__global__ void wgrad_pad_kernel(size_t n, size_t F, size_t Fp, const float* __restrict__ in, size_t ld, float* __restrict__ out) {
  const size_t total = n * Fp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Fp, c = i % Fp;
    out[i] = c < F ? __ldg(in + r * ld + c) : 0.f;
  }
}

inline bool tma_ok(const float* p, size_t ld) { return (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(p) % 16 == 0); }

}  // namespace

// C_0 = A_0^T·B_0 (and C_1 = A_1^T·B_1 with one operand shared), A_i [nrows x Kx_i] (lda), B_i [nrows x My_i] (ldb).
int gemm_tc_wgrad_cat(const WgradCat& q, int passes, cudaStream_t st) {
  const size_t nrows = q.nrows;
  if (nrows < 4096 || q.dual < 0 || q.dual > 2) return GAI_ERR_UNSUPPORTED;
  const int na = q.dual == 1 ? 2 : 1, nb = q.dual == 2 ? 2 : 1;
  for (int i = 0; i < na; i++) if (q.Kx[i] < 1 || q.Kx[i] > (na == 2 ? 128u : 256u)) return GAI_ERR_UNSUPPORTED;
  for (int i = 0; i < nb; i++) if (q.My[i] < 1 || q.My[i] > 256) return GAI_ERR_UNSUPPORTED;
  if (!encode_fn()) return GAI_ERR_UNSUPPORTED;
  WgArgs g;
  memset(&g, 0, sizeof(g));
  g.nrows = nrows; g.passes = passes;
  static const int debug_knobs = [] {
    const int k = getenv("GAI_TC_DEBUG") ? atoi(getenv("GAI_TC_DEBUG")) : 0;
    if (k) fprintf(stderr, "libgai_b200: GAI_TC_DEBUG=%d — timing experiment, dense-transform RESULTS ARE WRONG (tools/gemm_probe.py only)\n", k);
    return k;
  }();
  g.debug = debug_knobs;
  g.dual_a = na == 2;
  if (na == 2) {
    g.mt = 2;
    for (int i = 0; i < 2; i++) { g.rows_t[i] = (int)q.Kx[i]; g.nbox_at[i] = (int)((q.Kx[i] + 31) / 32); }
    g.Kx = (int)(q.Kx[0] + q.Kx[1]);
  } else {
    g.Kx = (int)q.Kx[0];
    g.mt = g.Kx > 128 ? 2 : 1;
    g.rows_t[0] = g.Kx > 128 ? 128 : g.Kx;
    g.rows_t[1] = g.Kx > 128 ? g.Kx - 128 : 0;
    g.nbox_at[0] = (g.rows_t[0] + 31) / 32;
    g.nbox_at[1] = (g.rows_t[1] + 31) / 32;
  }
  g.nbox_b0 = (int)((q.My[0] + 31) / 32);
  g.nbox_b = g.nbox_b0 + (nb == 2 ? (int)((q.My[1] + 31) / 32) : 0);
  if (g.nbox_b > 8) return GAI_ERR_UNSUPPORTED;
  g.my0 = (int)q.My[0];
  g.My = (int)(q.My[0] + (nb == 2 ? q.My[1] : 0));
  g.n_mma = g.nbox_b * 32;
  g.a_bytes = (uint32_t)g.mt * 4u * WG_BOX;
  g.b_bytes = (uint32_t)g.nbox_b * WG_BOX;
  g.stage_bytes = 2 * (g.a_bytes + g.b_bytes);
  g.stages = (int)((200u * 1024u) / g.stage_bytes);
  if (g.stages > 8) g.stages = 8;
  if (g.stages < 2) return GAI_ERR_UNSUPPORTED;

  const size_t total_kb = (nrows + WG_BK - 1) / WG_BK;
  size_t grid = total_kb < (size_t)sm_count() ? total_kb : (size_t)sm_count();
  g.blocks_per_cta = (total_kb + grid - 1) / grid;
  grid = (total_kb + g.blocks_per_cta - 1) / g.blocks_per_cta;  // every CTA owns at least one k-block
  g.my0p = (g.my0 + 3) / 4 * 4;
  g.ldp = g.my0p + ((g.My - g.my0) + 3) / 4 * 4;

  // workspace slot 0: per-CTA partials; slot 2: padded copies of operands TMA cannot address
  void* ws = nullptr;
  int rc = workspace(sizeof(float) * grid * g.Kx * g.ldp, &ws, st);
  if (rc != GAI_OK) return rc;
  g.partial = reinterpret_cast<float*>(ws);
  const float* src[4] = {q.A[0], na == 2 ? q.A[1] : q.A[0], q.B[0], nb == 2 ? q.B[1] : q.B[0]};
  size_t ld[4] = {q.lda[0], na == 2 ? q.lda[1] : q.lda[0], q.ldb[0], nb == 2 ? q.ldb[1] : q.ldb[0]};
  size_t cols[4] = {q.Kx[0], na == 2 ? q.Kx[1] : q.Kx[0], q.My[0], nb == 2 ? q.My[1] : q.My[0]};
  const bool used[4] = {true, na == 2, true, nb == 2};
  size_t pad_elems = 0;
  for (int i = 0; i < 4; i++)
    if (used[i] && !tma_ok(src[i], ld[i])) pad_elems += (nrows * ((cols[i] + 3) / 4 * 4) + 63) / 64 * 64;
  if (pad_elems) {
    void* ws2 = nullptr;
    rc = workspace_slot(2, sizeof(float) * pad_elems + 512, &ws2, st);
    if (rc != GAI_OK) return rc;
    float* p = reinterpret_cast<float*>(ws2);
    const size_t cap = (size_t)sm_count() * 32;
    for (int i = 0; i < 4; i++) {
      if (!used[i] || tma_ok(src[i], ld[i])) continue;
      const size_t cp = (cols[i] + 3) / 4 * 4;
      const size_t blocks = (nrows * cp + 255) / 256;
      wgrad_pad_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(nrows, cols[i], cp, src[i], ld[i], p);
      GAI_LAUNCH_CHECK();
      src[i] = p; ld[i] = cp; cols[i] = cp;
      p += (nrows * cp + 63) / 64 * 64;
    }
  }
  CUtensorMap maps[4];
  for (int i = 0; i < 4; i++) {
    if (!used[i]) { maps[i] = maps[i - 1]; continue; }
    if (!make_map_f32(&maps[i], src[i], nrows, cols[i], ld[i], WG_BK, true, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return set_error(GAI_ERR_CUDA, "gemm_tc_wgrad", "cuTensorMapEncodeTiled failed");
  }

  const size_t smem = (size_t)g.stages * g.stage_bytes + 1024;
  static bool configured = false;
  if (!configured) {
    GAI_CUDA(cudaFuncSetAttribute(gemm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    configured = true;
  }
  gemm_tc_wgrad_kernel<<<(unsigned)grid, WG_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], g);
  GAI_LAUNCH_CHECK();
  const int n = g.Kx * g.My;
  const int kx0 = na == 2 ? (int)q.Kx[0] : g.Kx, my0 = nb == 2 ? (int)q.My[0] : g.My;
  float* C1 = q.dual ? q.C[1] : q.C[0];
  const size_t ldc1 = q.dual ? q.ldc[1] : q.ldc[0];
  wgrad_reduce_kernel<<<(n * 8 + 255) / 256, 256, 0, st>>>(g.partial, q.C[0], q.ldc[0], C1, ldc1, g.Kx, g.My, g.ldp, kx0, my0, (int)grid, q.accum);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

// C[Kx x My] (+)= A^T · B,  A [nrows x Kx] (lda), B [nrows x My] (ldb).
int gemm_tc_wgrad(size_t Kx, size_t My, size_t nrows, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int accum,
                  int flags, int passes, cudaStream_t st) {
  if (Kx < 1 || Kx > 256 || My < 1 || My > 256 || flags != 0) return GAI_ERR_UNSUPPORTED;
  WgradCat q;
  q.nrows = nrows; q.dual = 0;
  q.A[0] = A; q.lda[0] = lda; q.Kx[0] = Kx; q.B[0] = B; q.ldb[0] = ldb; q.My[0] = My; q.C[0] = C; q.ldc[0] = ldc; q.accum = accum;
  return gemm_tc_wgrad_cat(q, passes, st);
}

}  // namespace gai
