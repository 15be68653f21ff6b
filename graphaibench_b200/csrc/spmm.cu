// Neighbour aggregation (CSR SpMM) for sm_100a: GCN symmetric-normalised, SAGE mean (forward / transposed) and
// explicit edge values (GAT), one kernel family.
//
// Replaces update_all_gcn / update_all_sage / reduce_warp+reduce_cta (include/gnn/graph_operations.h:8-178) and the
// CPU loops they mirror (src/gnn/gconv/gcn_aggregator.cpp:48-77, sage_aggregator.cpp:7-54, gat_aggregator.cpp:26-45).
//
// Design (B200: HBM/L2-latency bound gather; no tensor-core shape here):
//   * row-split by degree bucket. Rows with deg <= HUB_DEGREE: a group of G lanes (G = 4..32, chosen from the
//     feature width so that one 128-bit load per lane covers the row) owns one output row and keeps it in
//     registers; the group loads G column indices + edge weights with one coalesced request, broadcasts them by
//     shuffle, and issues U=4 independent 128-bit neighbour-row loads (ld.global.nc) before consuming them.
//   * hub rows (deg > HUB_DEGREE): one CTA per row. All 8 warps gather and scale neighbour rows into a shared-memory
//     tile in parallel; then one thread per column adds the tile's entries IN EDGE ORDER.
//   * numerics: acc = fadd_rn(acc, fmul_rn(w, x)) per edge, sequential per column — exactly the reference CPU
//     path's scale()+vadd() (math_functions.cpp:266,336), so results are bit-identical for every row length,
//     including hub rows (the parallel part is only the gather).
//   * fused: zero-init (no memset pass), optional "+ addend" and ReLU epilogue, leading dimensions (so the output can
//     land inside a wider buffer), row ranges (1D partition: interior vs boundary rows).
#include "gai_internal.cuh"

namespace {

enum Mode { M_GCN = 0, M_MEAN = 1, M_MEAN_T = 2, M_EDGE = 3, M_EDGE_PERM = 4 };

struct SpmmArgs {
  const uint32_t* rowptr;
  const uint32_t* colidx;
  const float* norm;
  const float* vals;
  const uint32_t* perm;
  const float* in;
  float* out;
  const float* addend;
  int F, ld_in, ld_out;
  uint32_t row_begin, row_end;
  int mode, flags;
  uint32_t hub_threshold;
};

template <int VEC> struct VecT;
template <> struct VecT<4> { using T = float4; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<1> { using T = float; };

template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float (&r)[VEC]) {
  if constexpr (VEC == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w; }
  else if constexpr (VEC == 2) { float2 t = __ldg(reinterpret_cast<const float2*>(p)); r[0] = t.x; r[1] = t.y; }
  else { r[0] = __ldg(p); }
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float (&r)[VEC]) {
  if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
  else if constexpr (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]);
  else *p = r[0];
}

__device__ __forceinline__ float edge_weight(const SpmmArgs& a, float wrow, uint32_t idx, uint32_t c) {
  switch (a.mode) {
    case M_GCN: return __fmul_rn(wrow, __ldg(a.norm + c));  // b = a_i * a_j (gcn_aggregator.cpp:66)
    case M_MEAN: return wrow;                               // 1/deg_i (sage_aggregator.cpp:17)
    case M_MEAN_T: return __ldg(a.norm + c);                // 1/deg_j (sage_aggregator.cpp:41)
    case M_EDGE: return __ldg(a.vals + idx);
    default: return __ldg(a.vals + __ldg(a.perm + idx));
  }
}

constexpr int U = 4;  // independent neighbour rows in flight per lane

template <int VEC, int G, int K>
__global__ void __launch_bounds__(256) spmm_rows_kernel(const SpmmArgs a) {
  constexpr int ROWS_PER_WARP = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int grp = lane / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const uint64_t warp_global = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t row64 = (uint64_t)a.row_begin + warp_global * ROWS_PER_WARP + grp;
  if (row64 >= a.row_end) return;
  const uint32_t row = (uint32_t)row64;
  const uint32_t s = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
  if (e - s > a.hub_threshold) return;
  const int nchunks = a.F / VEC;
  const float wrow = (a.mode == M_GCN || a.mode == M_MEAN) ? __ldg(a.norm + row) : 0.0f;

  for (int cb = 0; cb < nchunks; cb += G * K) {
    float acc[K][VEC];
    bool act[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      act[k] = (cb + gl + G * k) < nchunks;
#pragma unroll
      for (int v = 0; v < VEC; v++) acc[k][v] = 0.0f;
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
        float x[U][K][VEC];
        float ww[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int jj = j + u;
          const uint32_t cc = __shfl_sync(gmask, c, jj, G);
          ww[u] = __shfl_sync(gmask, w, jj, G);
          const float* src = a.in + (size_t)cc * a.ld_in + (size_t)(cb + gl) * VEC;
#pragma unroll
          for (int k = 0; k < K; k++) {
            if (jj < cnt && act[k]) ldv<VEC>(src + (size_t)G * k * VEC, x[u][k]);
            else {
#pragma unroll
              for (int v = 0; v < VEC; v++) x[u][k][v] = 0.0f;
            }
          }
          if (jj >= cnt) ww[u] = 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
          for (int k = 0; k < K; k++)
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[k][v] = __fadd_rn(acc[k][v], __fmul_rn(ww[u], x[u][k][v]));
      }
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (!act[k]) continue;
      const size_t col = (size_t)(cb + gl + G * k) * VEC;
      float r[VEC];
#pragma unroll
      for (int v = 0; v < VEC; v++) r[v] = acc[k][v];
      if (a.flags & GAI_EPI_ADD) {
        float ad[VEC];
        ldv<VEC>(a.addend + (size_t)row * a.ld_out + col, ad);
#pragma unroll
        for (int v = 0; v < VEC; v++) r[v] = __fadd_rn(r[v], ad[v]);
      }
      if (a.flags & GAI_EPI_RELU) {
#pragma unroll
        for (int v = 0; v < VEC; v++) r[v] = r[v] > 0.0f ? r[v] : 0.0f;
      }
      stv<VEC>(a.out + (size_t)row * a.ld_out + col, r);
    }
  }
}

// Hub rows: one CTA per row; parallel gather into shared memory, then an in-order add per column.
constexpr int HUB_THREADS = 256;
constexpr int HUB_KMAX = 4;  // columns per thread per column block (block = 1024 columns)

template <int VEC>
__global__ void __launch_bounds__(HUB_THREADS) spmm_hub_kernel(const SpmmArgs a, const uint32_t* __restrict__ hub_rows, int chunk_edges) {
  extern __shared__ float tile[];  // [chunk_edges][Fb]
  const uint32_t row = hub_rows[blockIdx.x];
  if (row < a.row_begin || row >= a.row_end) return;
  const uint32_t s = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float wrow = (a.mode == M_GCN || a.mode == M_MEAN) ? __ldg(a.norm + row) : 0.0f;
  for (int cb = 0; cb < a.F; cb += HUB_THREADS * HUB_KMAX) {
    const int Fb = (a.F - cb) < HUB_THREADS * HUB_KMAX ? (a.F - cb) : HUB_THREADS * HUB_KMAX;
    const int nch = Fb / VEC;
    float acc[HUB_KMAX];
#pragma unroll
    for (int k = 0; k < HUB_KMAX; k++) acc[k] = 0.0f;
    for (uint32_t base = s; base < e; base += chunk_edges) {
      const int cnt = (e - base) < (uint32_t)chunk_edges ? (int)(e - base) : chunk_edges;
      for (int j0 = warp * 2; j0 < cnt; j0 += (HUB_THREADS / 32) * 2) {
        uint32_t c[2];
        float w[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int j = j0 + u;
          c[u] = 0; w[u] = 0.0f;
          if (j < cnt) {
            c[u] = __ldg(a.colidx + base + j);
            w[u] = edge_weight(a, wrow, base + j, c[u]);
          }
        }
        for (int ch = lane; ch < nch; ch += 32) {
          float x[2][VEC];
#pragma unroll
          for (int u = 0; u < 2; u++) {
            if (j0 + u < cnt) ldv<VEC>(a.in + (size_t)c[u] * a.ld_in + cb + (size_t)ch * VEC, x[u]);
          }
#pragma unroll
          for (int u = 0; u < 2; u++) {
            if (j0 + u < cnt) {
#pragma unroll
              for (int v = 0; v < VEC; v++) tile[(size_t)(j0 + u) * Fb + ch * VEC + v] = __fmul_rn(w[u], x[u][v]);
            }
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < HUB_KMAX; k++) {
        const int col = threadIdx.x + k * HUB_THREADS;
        if (col < Fb) {
          float r = acc[k];
          for (int j = 0; j < cnt; j++) r = __fadd_rn(r, tile[(size_t)j * Fb + col]);
          acc[k] = r;
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < HUB_KMAX; k++) {
      const int col = threadIdx.x + k * HUB_THREADS;
      if (col < Fb) {
        float r = acc[k];
        const size_t o = (size_t)row * a.ld_out + cb + col;
        if (a.flags & GAI_EPI_ADD) r = __fadd_rn(r, __ldg(a.addend + o));
        if (a.flags & GAI_EPI_RELU) r = r > 0.0f ? r : 0.0f;
        a.out[o] = r;
      }
    }
  }
}

inline bool aligned(const void* p, size_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <int VEC>
int launch_rows(const SpmmArgs& a, cudaStream_t st) {
  const int nchunks = a.F / VEC;
  const uint64_t rows = (uint64_t)a.row_end - a.row_begin;
  int G = 4;
  while (G < 32 && G < nchunks) G <<= 1;
  int K = 1;
  if (G == 32) { K = (nchunks + 31) / 32; K = K <= 1 ? 1 : (K <= 2 ? 2 : 4); }
  const uint64_t rows_per_cta = (uint64_t)8 * (32 / G);
  const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
  if (grid == 0) return GAI_OK;
#define GAI_SPMM_CASE(g_, k_) spmm_rows_kernel<VEC, g_, k_><<<grid, 256, 0, st>>>(a)
  if (G == 4) GAI_SPMM_CASE(4, 1);
  else if (G == 8) GAI_SPMM_CASE(8, 1);
  else if (G == 16) GAI_SPMM_CASE(16, 1);
  else if (K == 1) GAI_SPMM_CASE(32, 1);
  else if (K == 2) GAI_SPMM_CASE(32, 2);
  else GAI_SPMM_CASE(32, 4);
#undef GAI_SPMM_CASE
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

template <int VEC>
int launch_hub(const SpmmArgs& a, const gai_csr* g, cudaStream_t st) {
  if (g->n_hub == 0) return GAI_OK;
  const int Fb = a.F < HUB_THREADS * HUB_KMAX ? a.F : HUB_THREADS * HUB_KMAX;
  int chunk = (int)((48 * 1024) / (sizeof(float) * (size_t)Fb));
  if (chunk > 64) chunk = 64;
  if (chunk < 1) chunk = 1;
  const size_t smem = sizeof(float) * (size_t)chunk * Fb;
  spmm_hub_kernel<VEC><<<g->n_hub, HUB_THREADS, smem, st>>>(a, g->hub_rows, chunk);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
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
  SpmmArgs a;
  a.rowptr = g->rowptr; a.colidx = g->colidx;
  a.norm = (mode == M_GCN) ? g->norm_gcn : g->norm_mean;
  a.vals = vals; a.perm = perm; a.in = in; a.out = out; a.addend = addend;
  a.F = F; a.ld_in = ld_in; a.ld_out = ld_out; a.row_begin = rb; a.row_end = re;
  a.mode = mode; a.flags = flags;
  a.hub_threshold = g->n_hub ? gai::HUB_DEGREE : 0xffffffffu;
  cudaStream_t st = gai::S(stream);
  const bool v4 = (F % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0) && aligned(in, 16) && aligned(out, 16) && aligned(addend, 16);
  const bool v2 = (F % 2 == 0) && (ld_in % 2 == 0) && (ld_out % 2 == 0) && aligned(in, 8) && aligned(out, 8) && aligned(addend, 8);
  int rc;
  if (v4) { rc = launch_rows<4>(a, st); if (rc == GAI_OK) rc = launch_hub<4>(a, g, st); }
  else if (v2) { rc = launch_rows<2>(a, st); if (rc == GAI_OK) rc = launch_hub<2>(a, g, st); }
  else { rc = launch_rows<1>(a, st); if (rc == GAI_OK) rc = launch_hub<1>(a, g, st); }
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
