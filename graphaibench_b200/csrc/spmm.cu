// Neighbour aggregation (CSR SpMM) for sm_100a: GCN symmetric-normalised, SAGE mean (forward / transposed) and
// explicit edge values (GAT), one kernel family.
//
// Replaces update_all_gcn / update_all_sage / reduce_warp+reduce_cta (include/gnn/graph_operations.h:8-178) and the
// CPU loops they mirror (src/gnn/gconv/gcn_aggregator.cpp:48-77, sage_aggregator.cpp:7-54, gat_aggregator.cpp:26-45).
//
// Design (B200: an HBM/L2 gather; no tensor-core shape here):
//   * every input row is read with 128-bit loads: if F % 4 != 0 or the layout is misaligned, the input is first copied
//     into a zero-padded workspace with ld = ceil4(F) (one N x F pass, ~1% of the gather traffic).
//   * row-split by degree bucket.
//       light rows (deg <= hub_degree): a group of G lanes (G = 4..32, from the feature width) owns one output row in
//         registers; the group loads G column indices + edge weights with one coalesced request, broadcasts them by
//         shuffle, and keeps U independent 128-bit neighbour-row loads in flight per lane.
//       hub rows (deg > hub_degree, a per-graph threshold = a warp's fair share of the edges): one warp-specialised CTA
//         per row. 16 producer warps gather + scale neighbour rows into a shared-memory ring (one slot per producer,
//         mbarrier full/empty pairs); 4 consumer warps (one thread per 4 columns) add the staged products IN EDGE
//         ORDER. The gather runs at SM bandwidth while the add chain stays sequential.
//   * numerics: acc = fadd_rn(acc, fmul_rn(w, x)) per edge, sequential per column — exactly the reference CPU path's
//     scale()+vadd() (math_functions.cpp:266,336): results are bit-identical for every row length, hub rows included.
//   * fused: zero-init (no memset pass), optional "+ addend" and ReLU epilogue, leading dimensions, row ranges
//     (1D partition: interior vs boundary rows).
#include "gai_internal.cuh"

namespace {

enum Mode { M_GCN = 0, M_MEAN = 1, M_MEAN_T = 2, M_EDGE = 3, M_EDGE_PERM = 4 };

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
  unsigned long long scramble;  // odd multiplier coprime to the row count (1 = natural order)
};

__device__ __forceinline__ float edge_weight(const SpmmArgs& a, float wrow, uint32_t idx, uint32_t c) {
  switch (a.mode) {
    case M_GCN: return __fmul_rn(wrow, __ldg(a.norm + c));  // b = a_i * a_j (gcn_aggregator.cpp:66)
    case M_MEAN: return wrow;                               // 1/deg_i (sage_aggregator.cpp:17)
    case M_MEAN_T: return __ldg(a.norm + c);                // 1/deg_j (sage_aggregator.cpp:41)
    case M_EDGE: return __ldg(a.vals + idx);
    default: return __ldg(a.vals + __ldg(a.perm + idx));
  }
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
    *reinterpret_cast<float4*>(a.out + o) = r;
  } else {
    const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (chunk * 4 + k < a.F) {
        float t = v[k];
        if (a.flags & GAI_EPI_ADD) t = __fadd_rn(t, a.addend[o + k]);
        if (a.flags & GAI_EPI_RELU) t = t > 0.f ? t : 0.f;
        a.out[o + k] = t;
      }
    }
  }
}

// ---- light rows -----------------------------------------------------------------------------------------------------
// One output row per group of G lanes, kept in registers.
template <int G, int K>
__device__ __forceinline__ void spmm_one_row(const SpmmArgs& a, uint32_t row, int gl, unsigned gmask) {
  constexpr int UMAX = (K == 1) ? 8 : (K == 2 ? 4 : 2);
  constexpr int U = G < UMAX ? G : UMAX;  // independent neighbour rows in flight per lane (up to 8 float4)
  const uint32_t s = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
  if (e - s > a.hub_threshold) return;
  const float wrow = (a.mode == M_GCN || a.mode == M_MEAN) ? __ldg(a.norm + row) : 0.0f;
  const float4* in4 = reinterpret_cast<const float4*>(a.in);
  const size_t ld4 = (size_t)a.ld_in >> 2;

  for (int cb = 0; cb < a.nchunks; cb += G * K) {
    float4 acc[K];
    bool act[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      act[k] = (cb + gl + G * k) < a.nchunks;
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (uint32_t base = s; base < e; base += G) {
      const uint32_t idx = base + gl;
      uint32_t c = 0;
      float w = 0.0f;
      if (idx < e) {
        c = __ldg(a.colidx + idx);
        w = edge_weight(a, wrow, idx, c);
      }
      const int cnt = (e - base) < (uint32_t)G ? (int)(e - base) : G;
#pragma unroll
      for (int j = 0; j < G; j += U) {
        if (j >= cnt) break;
        float4 x[U][K];
        float ww[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int jj = j + u;
          const uint32_t cc = __shfl_sync(gmask, c, jj, G);
          ww[u] = __shfl_sync(gmask, w, jj, G);
          const float4* src = in4 + (size_t)cc * ld4 + (cb + gl);
#pragma unroll
          for (int k = 0; k < K; k++) {
            if (jj < cnt && act[k]) x[u][k] = __ldg(src + G * k);
            else x[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (jj >= cnt) ww[u] = 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
          for (int k = 0; k < K; k++) {
            acc[k].x = __fadd_rn(acc[k].x, __fmul_rn(ww[u], x[u][k].x));
            acc[k].y = __fadd_rn(acc[k].y, __fmul_rn(ww[u], x[u][k].y));
            acc[k].z = __fadd_rn(acc[k].z, __fmul_rn(ww[u], x[u][k].z));
            acc[k].w = __fadd_rn(acc[k].w, __fmul_rn(ww[u], x[u][k].w));
          }
      }
    }
#pragma unroll
    for (int k = 0; k < K; k++)
      if (act[k]) store_chunk(a, row, cb + gl + G * k, acc[k]);
  }
}

// Persistent warps (grid = 4 CTAs per SM): each warp claims ROW_CHUNK consecutive row-groups at a time from a global
// counter, so a warp that drew a long row does not hold idle siblings resident (power-law graphs: 40% of the rows of the
// bench graph are empty while others are 8 K edges long).
constexpr int ROW_CHUNK = 8;

template <int G, int K>
__global__ void __launch_bounds__(256, 4) spmm_rows_kernel(const SpmmArgs a, unsigned long long* __restrict__ counter) {
  constexpr int ROWS_PER_WARP = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int grp = lane / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const unsigned long long nrows = (unsigned long long)a.row_end - a.row_begin;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)(ROW_CHUNK * ROWS_PER_WARP));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= nrows) break;
#pragma unroll 1
    for (int i = 0; i < ROW_CHUNK; i++) {
      const unsigned long long q = base + (unsigned long long)i * ROWS_PER_WARP + grp;
      // multiplicative permutation of the work order (scramble is coprime to nrows): long rows that sit next to each
      // other in id space (R-MAT, BFS orderings) land in different chunks; every row is still produced exactly once
      if (q < nrows) spmm_one_row<G, K>(a, (uint32_t)(a.row_begin + (q * a.scramble) % nrows), gl, gmask);
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

// ES = edges per stage (32, 16, 8 or 4). Dynamic smem: HUB_PROD_WARPS * ES * min(nchunks,128) float4.
template <int ES>
__global__ void __launch_bounds__(HUB_THREADS) spmm_hub_kernel(const SpmmArgs a, const uint32_t* __restrict__ hub_rows) {
  extern __shared__ float4 ring[];
  __shared__ uint64_t full_bar[HUB_PROD_WARPS], empty_bar[HUB_PROD_WARPS];
  const uint32_t row = hub_rows[blockIdx.x];
  if (row < a.row_begin || row >= a.row_end) return;  // uniform for the CTA
  const uint32_t s = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
  const uint32_t nstages = (e - s + ES - 1) / ES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4* in4 = reinterpret_cast<const float4*>(a.in);
  const size_t ld4 = (size_t)a.ld_in >> 2;
  const int nch_max = a.nchunks < HUB_MAX_CHUNKS ? a.nchunks : HUB_MAX_CHUNKS;
  const size_t slot_stride = (size_t)ES * nch_max;
  uint32_t blk = 0;  // column-block index; slot p has been used blk * uses(p) times before this block (mbarrier phase bookkeeping)

  if (threadIdx.x == 0) {
    for (int i = 0; i < HUB_PROD_WARPS; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], HUB_CONS_WARPS); }
  }
  __syncthreads();

  for (int cb = 0; cb < a.nchunks; cb += HUB_MAX_CHUNKS) {
    const int nch = (a.nchunks - cb) < HUB_MAX_CHUNKS ? (a.nchunks - cb) : HUB_MAX_CHUNKS;
    if (warp >= HUB_CONS_WARPS) {
      // ---------------- producers ----------------
      const int pw = warp - HUB_CONS_WARPS;
      const float wrow = (a.mode == M_GCN || a.mode == M_MEAN) ? __ldg(a.norm + row) : 0.0f;
      float4* slot = ring + (size_t)pw * slot_stride;
      const uint32_t round0 = blk * ((nstages + HUB_PROD_WARPS - 1 - pw) / HUB_PROD_WARPS);
      // prefetch the first stage's indices / weights (lane l < ES holds edge l of the stage)
      uint32_t c = 0; float w = 0.0f;
      {
        const uint32_t idx = s + (uint32_t)pw * ES + lane;
        if (lane < ES && idx < e) { c = __ldg(a.colidx + idx); w = edge_weight(a, wrow, idx, c); }
      }
      for (uint32_t k = pw, r = 0; k < nstages; k += HUB_PROD_WARPS, r++) {
        const uint32_t base = s + k * ES;
        const int cnt = (e - base) < (uint32_t)ES ? (int)(e - base) : ES;
        const uint32_t cur_c = c; const float cur_w = w;
        // prefetch the next stage this warp owns
        c = 0; w = 0.0f;
        {
          const uint64_t nidx = (uint64_t)base + (uint64_t)HUB_PROD_WARPS * ES + lane;
          if (lane < ES && nidx < e) { c = __ldg(a.colidx + nidx); w = edge_weight(a, wrow, (uint32_t)nidx, c); }
        }
        mbar_wait(&empty_bar[pw], ((round0 + r) & 1) ^ 1);
        for (int ch0 = 0; ch0 < nch; ch0 += 32) {
          const int ch = ch0 + lane;
          const bool chv = ch < nch;
          constexpr int UB = ES < 8 ? ES : 8;
#pragma unroll
          for (int j0 = 0; j0 < ES; j0 += UB) {
            float4 x[UB]; float ww[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
              const uint32_t cc = __shfl_sync(0xffffffffu, cur_c, j0 + u);
              ww[u] = __shfl_sync(0xffffffffu, cur_w, j0 + u);
              if (chv && j0 + u < cnt) x[u] = __ldg(in4 + (size_t)cc * ld4 + cb + ch);
            }
#pragma unroll
            for (int u = 0; u < UB; u++) {
              if (chv && j0 + u < cnt) {
                float4 p;
                p.x = __fmul_rn(ww[u], x[u].x); p.y = __fmul_rn(ww[u], x[u].y); p.z = __fmul_rn(ww[u], x[u].z); p.w = __fmul_rn(ww[u], x[u].w);
                slot[(size_t)(j0 + u) * nch + ch] = p;
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

int launch_rows(const SpmmArgs& a, const gai_csr* g, cudaStream_t st) {
  const uint64_t rows = (uint64_t)a.row_end - a.row_begin;
  int G = 4;
  while (G < 32 && G < a.nchunks) G <<= 1;
  int K = 1;
  if (G == 32) { K = (a.nchunks + 31) / 32; K = K <= 1 ? 1 : (K <= 2 ? 2 : 4); }
  const uint64_t rows_per_fetch = (uint64_t)ROW_CHUNK * (32 / G);
  const uint64_t fetches = (rows + rows_per_fetch - 1) / rows_per_fetch;
  uint64_t ctas = (fetches + 7) / 8;
  const uint64_t persistent = (uint64_t)gai::sm_count() * 4;
  if (ctas > persistent) ctas = persistent;
  if (ctas == 0) return GAI_OK;
  const unsigned grid = (unsigned)ctas;
  SpmmArgs b = a;
  b.scramble = 1;
  if (rows > 64) {
    static const unsigned long long primes[] = {1000003ull, 998244353ull, 2654435761ull, 40503ull, 7919ull};
    for (unsigned long long pr : primes) {
      unsigned long long x = pr % rows, y = rows;
      while (y) { const unsigned long long t = x % y; x = y; y = t; }  // gcd(pr mod rows, rows)
      if (x == 1 && (pr % rows) > 1) { b.scramble = pr % rows; break; }
    }
  }
  // rotating work counters: launches on one stream are ordered; the rotation keeps up to 16 launches that overlap on
  // different streams (interior / boundary rows of the 1D partition) from sharing a counter
  unsigned long long* ctr = g->row_counters + (__atomic_fetch_add(&const_cast<gai_csr*>(g)->counter_seq, 1u, __ATOMIC_RELAXED) % 16u);
  GAI_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
  if (G == 4) spmm_rows_kernel<4, 1><<<grid, 256, 0, st>>>(b, ctr);
  else if (G == 8) spmm_rows_kernel<8, 1><<<grid, 256, 0, st>>>(b, ctr);
  else if (G == 16) spmm_rows_kernel<16, 1><<<grid, 256, 0, st>>>(b, ctr);
  else if (K == 1) spmm_rows_kernel<32, 1><<<grid, 256, 0, st>>>(b, ctr);
  else if (K == 2) spmm_rows_kernel<32, 2><<<grid, 256, 0, st>>>(b, ctr);
  else spmm_rows_kernel<32, 4><<<grid, 256, 0, st>>>(b, ctr);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

template <int ES>
int launch_hub_es(const SpmmArgs& a, const gai_csr* g, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    GAI_CUDA(cudaFuncSetAttribute(spmm_hub_kernel<ES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    configured = true;
  }
  spmm_hub_kernel<ES><<<g->n_hub, HUB_THREADS, smem, st>>>(a, g->hub_rows);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

int launch_hub(const SpmmArgs& a, const gai_csr* g, cudaStream_t st) {
  if (g->n_hub == 0) return GAI_OK;
  const int nch = a.nchunks < HUB_MAX_CHUNKS ? a.nchunks : HUB_MAX_CHUNKS;
  // largest stage size in {32, 16, 8, 4} edges whose 16-slot ring fits 200 KB of shared memory (longer stages amortise the
  // consumer's per-stage barrier round trip over more in-order adds)
  auto bytes = [&](int es) { return (size_t)HUB_PROD_WARPS * es * nch * sizeof(float4); };
  if (bytes(32) <= 200 * 1024) return launch_hub_es<32>(a, g, bytes(32), st);
  if (bytes(16) <= 192 * 1024) return launch_hub_es<16>(a, g, bytes(16), st);
  if (bytes(8) <= 192 * 1024) return launch_hub_es<8>(a, g, bytes(8), st);
  return launch_hub_es<4>(a, g, bytes(4), st);
}

int spmm_dispatch(gai_csr_t g, int mode, uint32_t rb, uint32_t re, int F, const float* vals, const uint32_t* perm, const float* in,
                  int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr);
  GAI_CHECK_ARG(rb <= re && re <= g->nv);
  if (re == rb) return GAI_OK;  // empty graph / empty row range: nothing to do (buffers may be NULL)
  GAI_CHECK_ARG(in != nullptr && out != nullptr);
  GAI_CHECK_ARG(F > 0 && ld_in >= F && ld_out >= F);
  GAI_CHECK_ARG(!(flags & GAI_EPI_ADD) || addend != nullptr);
  GAI_CHECK_ARG(mode < M_EDGE || vals != nullptr);
  GAI_CHECK_ARG(in != out);
  cudaStream_t st = gai::S(stream);
  SpmmArgs a;
  a.rowptr = g->rowptr; a.colidx = g->colidx;
  a.norm = (mode == M_GCN) ? g->norm_gcn : g->norm_mean;
  a.vals = vals; a.perm = perm; a.out = out; a.addend = addend;
  a.F = F; a.nchunks = (F + 3) / 4; a.ld_out = ld_out; a.row_begin = rb; a.row_end = re;
  a.mode = mode; a.flags = flags;
  a.hub_threshold = g->n_hub ? g->hub_degree : 0xffffffffu;
  a.out_vec = (F % 4 == 0) && (ld_out % 4 == 0) && aligned16(out) && aligned16(addend);
  if ((F % 4 == 0) && (ld_in % 4 == 0) && aligned16(in)) {
    a.in = in; a.ld_in = ld_in;
  } else {
    // gather source must be 128-bit loadable: stage a zero-padded copy (all nv rows can be neighbours)
    const int Fp = a.nchunks * 4;
    void* ws = nullptr;
    int rc = gai::workspace_slot(1, sizeof(float) * (size_t)g->nv * Fp, &ws);
    if (rc != GAI_OK) return rc;
    const size_t total = (size_t)g->nv * Fp;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)gai::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    pad_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(g->nv, F, Fp, in, ld_in, reinterpret_cast<float*>(ws));
    GAI_LAUNCH_CHECK();
    a.in = reinterpret_cast<const float*>(ws); a.ld_in = Fp;
  }
  // Hub rows go first, on a high-priority side stream (fork/join with events): their CTAs are the long poles (one
  // 94 K-edge row is ~0.3 ms of in-order adds), the persistent light-row warps on `st` fill the other SMs meanwhile.
  int rc = GAI_OK;
  if (g->n_hub) {
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
  if (g->n_hub) GAI_CUDA(cudaStreamWaitEvent(st, g->ev_join, 0));
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
int gai_spmm_gcn_rows(gai_csr_t g, uint32_t rb, uint32_t re, int F, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream) {
  return spmm_dispatch(g, M_GCN, rb, re, F, nullptr, nullptr, in, ld_in, out, ld_out, flags, addend, stream);
}
int gai_spmm_mean_rows(gai_csr_t g, uint32_t rb, uint32_t re, int F, const float* in, int ld_in, float* out, int ld_out, int transposed, int flags, const float* addend, gai_stream_t stream) {
  return spmm_dispatch(g, transposed ? M_MEAN_T : M_MEAN, rb, re, F, nullptr, nullptr, in, ld_in, out, ld_out, flags, addend, stream);
}

}  // extern "C"
