// GAT attention for sm_100a: scores + edge-softmax (forward); SDDMM, softmax/LeakyReLU backward, attention-vector
// gradients and the transposed aggregation (backward).
// Replaces compute_attn_score_warp / compute_scores_grad_warp / compute_alpha_grad_warp / csr2csc
// (include/gnn/graph_operations.h:190-467, src/gnn/gconv/gat_aggregator.cu:7-115) and restates the CPU path
// (src/gnn/gconv/gat_aggregator.cpp:57-200) with these changes of schedule, not of math:
//   * el_i = <alpha_l, z_i>, er_i = <alpha_r, z_i> are computed once per VERTEX (the reference recomputes
//     <alpha_r, z_j> once per EDGE, gat_aggregator.cpp:70);
//   * d_alpha_r = sum_e ds_e z_{dst(e)} is regrouped as Z^T·colsum with colsum_j = sum of ds over the edges that
//     point at j (read through the cached e->e^T permutation), d_alpha_l = Z^T·rowsum: one pass over Z instead of
//     an nnz x F gather, reduced in two deterministic stages (the reference GPU kernel uses float atomics);
//   * the transposed attention matrix is never materialised: the final SpMM reads vals[perm[e]].
// The per-edge aggregation itself goes through the bit-exact SpMM family in spmm.cu.
#include "gai_internal.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Per-head reductions over the lanes of a warp: with H heads (a power of two <= 32) lane l works on head l % H, so a butterfly over the
// offsets 16 .. H leaves every lane with the total of its own head. H = 1: the plain warp reduction.
__device__ __forceinline__ float head_sum(float v, int H) {
  for (int o = 16; o >= H; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float head_max(float v, int H) {
  for (int o = 16; o >= H; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Row-cooperative per-head reductions: a warp (light rows) or a whole CTA (hub rows; red holds 32 floats per warp).
template <bool CTA>
__device__ __forceinline__ float coop_hsum(float v, int H, float* red) {
  v = head_sum(v, H);
  if (!CTA) return v;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane < H) red[w * 32 + lane] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; i++) t += red[i * 32 + (lane & (H - 1))];
  return t;
}
template <bool CTA>
__device__ __forceinline__ float coop_hmax(float v, int H, float* red) {
  v = head_max(v, H);
  if (!CTA) return v;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane < H) red[w * 32 + lane] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < nw; i++) t = fmaxf(t, red[i * 32 + (lane & (H - 1))]);
  return t;
}

struct RowSel {
  const uint32_t* rowptr;
  const uint32_t* hub_rows;
  const uint32_t* order;   // rows by degree, longest first (hub rows at its head); NULL: natural order
  uint32_t nv, n_hub, hub_threshold;
};

// Resolves which row this warp / CTA owns; returns false if none.
template <bool CTA>
__device__ __forceinline__ bool pick_row(const RowSel& r, uint32_t& row, uint32_t& s, uint32_t& e, int& tid, int& nthr) {
  if (CTA) {
    if (blockIdx.x >= r.n_hub) return false;
    row = r.hub_rows[blockIdx.x];
    tid = threadIdx.x; nthr = blockDim.x;
  } else {
    // warp-per-row kernels walk the rows in DEGREE order: the eight warps of a CTA then own rows of (nearly) equal length. In natural
    // order a CTA stays resident until its longest row is done with one warp active — on the Reddit-shaped graph that held the score
    // kernels at a quarter of an item per clock and SM (13 ms per call with 8 heads).
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= r.nv) return false;
    row = r.order ? __ldg(r.order + w) : (uint32_t)w;
    tid = threadIdx.x & 31; nthr = 32;
  }
  s = __ldg(r.rowptr + row); e = __ldg(r.rowptr + row + 1);
  if (!CTA && (e - s) > r.hub_threshold) return false;
  return true;
}

// el/er per vertex (and head): one warp per row, coalesced. H heads of D = F / H columns each: el[i * H + h] = <alpha_l[hD..], z_i[hD..]>.
__global__ void el_er_kernel(uint32_t nv, int F, int H, const float* __restrict__ z, size_t ld, const float* __restrict__ al, const float* __restrict__ ar,
                             float* __restrict__ el, float* __restrict__ er) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nv) return;
  const float* x = z + (size_t)w * ld;
  const int D = F / H;
  for (int h = 0; h < H; h++) {
    float a = 0.f, b = 0.f;
    for (int k = h * D + lane; k < (h + 1) * D; k += 32) { const float v = x[k]; a += __ldg(al + k) * v; b += __ldg(ar + k) * v; }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { el[w * H + h] = a; er[w * H + h] = b; }
  }
}

// temp_scores[e] = el_i + er_j; norm_scores = softmax_row(LeakyReLU(temp_scores))   (gat_aggregator.cpp:62-77)
// Score arrays are edge-major with H entries per edge: item i = e * H + h. The items of a row are the contiguous range [s*H, e*H); thread
// `tid` walks them with a stride that is a multiple of H, so it only ever sees head tid % H and the per-head reductions are butterflies.
template <bool CTA>
__global__ void scores_kernel(const RowSel r, int H, const uint32_t* __restrict__ colidx, const float* __restrict__ el, const float* __restrict__ er,
                              float slope, float* __restrict__ temp_scores, float* __restrict__ norm_scores) {
  __shared__ float red[32 * 8];
  uint32_t row, s, e; int tid, nthr;
  if (!pick_row<CTA>(r, row, s, e, tid, nthr)) return;
  const int h = tid & (H - 1), hs = __ffs(H) - 1;
  const uint64_t i0 = (uint64_t)s * H, i1 = (uint64_t)e * H;
  const float eli = __ldg(el + (size_t)row * H + h);
  float mx = -INFINITY;
  // four independent (index -> er) load chains per thread and iteration
  uint64_t i = i0 + tid;
  for (; i + 3ull * nthr < i1; i += 4ull * nthr) {
    uint32_t c[4]; float t[4];
#pragma unroll
    for (int u = 0; u < 4; u++) c[u] = __ldg(colidx + ((i + (uint64_t)u * nthr) >> hs));
#pragma unroll
    for (int u = 0; u < 4; u++) t[u] = eli + __ldg(er + (size_t)c[u] * H + h);
#pragma unroll
    for (int u = 0; u < 4; u++) { temp_scores[i + (uint64_t)u * nthr] = t[u]; mx = fmaxf(mx, t[u] > 0.f ? t[u] : slope * t[u]); }
  }
  for (; i < i1; i += nthr) {
    const float t = eli + __ldg(er + (size_t)__ldg(colidx + (i >> hs)) * H + h);
    temp_scores[i] = t;
    mx = fmaxf(mx, t > 0.f ? t : slope * t);
  }
  mx = coop_hmax<CTA>(mx, H, red);
  // the thread re-reads only what it wrote itself: no barrier between the passes. The unnormalised exponentials are stored once and
  // rescaled in place (the reference's softmax: exp, sum, divide — math_functions.cpp:485-494).
  float sum = 0.f;
#pragma unroll 4
  for (uint64_t k = i0 + tid; k < i1; k += nthr) {
    const float t = temp_scores[k];
    const float p = expf((t > 0.f ? t : slope * t) - mx);
    norm_scores[k] = p;
    sum += p;
  }
  sum = coop_hsum<CTA>(sum, H, red);
#pragma unroll 4
  for (uint64_t k = i0 + tid; k < i1; k += nthr) norm_scores[k] = norm_scores[k] / sum;
}

// ---- 128-bit forms of the three per-(edge, head) passes for H % 4 == 0 ----------------------------------------------------------------
// A thread owns a float4 of four consecutive heads of one edge: 16 bytes per load instead of 4 quadruple the bytes a warp keeps in
// flight, which is what bounds these streaming passes (one warp per row: with 4-byte items the scalar kernels moved 1.4 TB/s on the
// Reddit-shaped graph with 8 heads). G4 = H / 4 float4 groups per edge; thread tid always sees group tid % G4, so the per-head reductions
// are butterflies over the offsets 16 .. G4 applied to each component.
__device__ __forceinline__ float4 f4_set(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float4 head_sum4(float4 v, int G4) {
  for (int o = 16; o >= G4; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  return v;
}
__device__ __forceinline__ float4 head_max4(float4 v, int G4) {
  for (int o = 16; o >= G4; o >>= 1) {
    v.x = fmaxf(v.x, __shfl_xor_sync(0xffffffffu, v.x, o)); v.y = fmaxf(v.y, __shfl_xor_sync(0xffffffffu, v.y, o));
    v.z = fmaxf(v.z, __shfl_xor_sync(0xffffffffu, v.z, o)); v.w = fmaxf(v.w, __shfl_xor_sync(0xffffffffu, v.w, o));
  }
  return v;
}
template <bool CTA, bool MAX>
__device__ __forceinline__ float4 coop_h4(float4 v, int G4, float4* red) {
  v = MAX ? head_max4(v, G4) : head_sum4(v, G4);
  if (!CTA) return v;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane < G4) red[w * 8 + lane] = v;
  __syncthreads();
  float4 t = MAX ? f4_set(-INFINITY) : f4_set(0.f);
  for (int i = 0; i < nw; i++) {
    const float4 q = red[i * 8 + (lane & (G4 - 1))];
    if (MAX) { t.x = fmaxf(t.x, q.x); t.y = fmaxf(t.y, q.y); t.z = fmaxf(t.z, q.z); t.w = fmaxf(t.w, q.w); }
    else { t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
  }
  return t;
}
__device__ __forceinline__ float lrelu(float t, float slope) { return t > 0.f ? t : slope * t; }

template <bool CTA>
__global__ void scores_kernel_v4(const RowSel r, int H, const uint32_t* __restrict__ colidx, const float* __restrict__ el, const float* __restrict__ er,
                                 float slope, float* __restrict__ temp_scores, float* __restrict__ norm_scores) {
  __shared__ float4 red[8 * 8];
  uint32_t row, s, e; int tid, nthr;
  if (!pick_row<CTA>(r, row, s, e, tid, nthr)) return;
  const int G4 = H >> 2, gs = __ffs(G4) - 1, hg = tid & (G4 - 1);
  const uint64_t q0 = (uint64_t)s * G4, q1 = (uint64_t)e * G4;
  float4* t4 = reinterpret_cast<float4*>(temp_scores);
  float4* n4 = reinterpret_cast<float4*>(norm_scores);
  const float4* er4 = reinterpret_cast<const float4*>(er);
  const float4 eli = __ldg(reinterpret_cast<const float4*>(el) + (size_t)row * G4 + hg);
  float4 mx = f4_set(-INFINITY);
  uint64_t q = q0 + tid;
  for (; q + 3ull * nthr < q1; q += 4ull * nthr) {
    uint32_t c[4]; float4 t[4];
#pragma unroll
    for (int u = 0; u < 4; u++) c[u] = __ldg(colidx + ((q + (uint64_t)u * nthr) >> gs));
#pragma unroll
    for (int u = 0; u < 4; u++) t[u] = __ldg(er4 + (size_t)c[u] * G4 + hg);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      t[u].x += eli.x; t[u].y += eli.y; t[u].z += eli.z; t[u].w += eli.w;
      t4[q + (uint64_t)u * nthr] = t[u];
      mx.x = fmaxf(mx.x, lrelu(t[u].x, slope)); mx.y = fmaxf(mx.y, lrelu(t[u].y, slope));
      mx.z = fmaxf(mx.z, lrelu(t[u].z, slope)); mx.w = fmaxf(mx.w, lrelu(t[u].w, slope));
    }
  }
  for (; q < q1; q += nthr) {
    float4 t = __ldg(er4 + (size_t)__ldg(colidx + (q >> gs)) * G4 + hg);
    t.x += eli.x; t.y += eli.y; t.z += eli.z; t.w += eli.w;
    t4[q] = t;
    mx.x = fmaxf(mx.x, lrelu(t.x, slope)); mx.y = fmaxf(mx.y, lrelu(t.y, slope));
    mx.z = fmaxf(mx.z, lrelu(t.z, slope)); mx.w = fmaxf(mx.w, lrelu(t.w, slope));
  }
  mx = coop_h4<CTA, true>(mx, G4, red);
  float4 sum = f4_set(0.f);
#pragma unroll 4
  for (uint64_t k = q0 + tid; k < q1; k += nthr) {
    const float4 t = t4[k];
    float4 p;
    p.x = expf(lrelu(t.x, slope) - mx.x); p.y = expf(lrelu(t.y, slope) - mx.y);
    p.z = expf(lrelu(t.z, slope) - mx.z); p.w = expf(lrelu(t.w, slope) - mx.w);
    n4[k] = p;
    sum.x += p.x; sum.y += p.y; sum.z += p.z; sum.w += p.w;
  }
  sum = coop_h4<CTA, false>(sum, G4, red);
#pragma unroll 4
  for (uint64_t k = q0 + tid; k < q1; k += nthr) {
    float4 p = n4[k];
    p.x = p.x / sum.x; p.y = p.y / sum.y; p.z = p.z / sum.z; p.w = p.w / sum.w;
    n4[k] = p;
  }
}

template <bool CTA>
__global__ void softmax_bwd_kernel_v4(const RowSel r, int H, float slope, const float* __restrict__ temp_scores, const float* __restrict__ p,
                                      float* __restrict__ ds, float* __restrict__ rowsum) {
  __shared__ float4 red[8 * 8];
  uint32_t row, s, e; int tid, nthr;
  if (!pick_row<CTA>(r, row, s, e, tid, nthr)) return;
  const int G4 = H >> 2;
  const uint64_t q0 = (uint64_t)s * G4, q1 = (uint64_t)e * G4;
  const float4* t4 = reinterpret_cast<const float4*>(temp_scores);
  const float4* p4 = reinterpret_cast<const float4*>(p);
  float4* d4 = reinterpret_cast<float4*>(ds);
  float4 dot = f4_set(0.f);
#pragma unroll 4
  for (uint64_t k = q0 + tid; k < q1; k += nthr) {
    const float4 a = p4[k], b = d4[k];
    dot.x += a.x * b.x; dot.y += a.y * b.y; dot.z += a.z * b.z; dot.w += a.w * b.w;
  }
  dot = coop_h4<CTA, false>(dot, G4, red);
  float4 rs = f4_set(0.f);
#pragma unroll 4
  for (uint64_t k = q0 + tid; k < q1; k += nthr) {
    const float4 a = p4[k], b = d4[k], t = t4[k];
    float4 v;
    v.x = a.x * (b.x - dot.x) * (t.x > 0.f ? 1.0f : slope); v.y = a.y * (b.y - dot.y) * (t.y > 0.f ? 1.0f : slope);
    v.z = a.z * (b.z - dot.z) * (t.z > 0.f ? 1.0f : slope); v.w = a.w * (b.w - dot.w) * (t.w > 0.f ? 1.0f : slope);
    d4[k] = v;
    rs.x += v.x; rs.y += v.y; rs.z += v.z; rs.w += v.w;
  }
  rs = coop_h4<CTA, false>(rs, G4, red);
  if (tid < G4) reinterpret_cast<float4*>(rowsum)[(size_t)row * G4 + tid] = rs;
}

template <bool CTA>
__global__ void colsum_kernel_v4(const RowSel r, int H, const uint32_t* __restrict__ perm, const float* __restrict__ ds, float* __restrict__ colsum) {
  __shared__ float4 red[8 * 8];
  uint32_t row, s, e; int tid, nthr;
  if (!pick_row<CTA>(r, row, s, e, tid, nthr)) return;
  const int G4 = H >> 2, gs = __ffs(G4) - 1, hg = tid & (G4 - 1);
  const uint64_t q0 = (uint64_t)s * G4, q1 = (uint64_t)e * G4;
  const float4* d4 = reinterpret_cast<const float4*>(ds);
  float4 cs = f4_set(0.f);
#pragma unroll 4
  for (uint64_t k = q0 + tid; k < q1; k += nthr) {
    const float4 v = __ldg(d4 + (size_t)__ldg(perm + (k >> gs)) * G4 + hg);
    cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
  }
  cs = coop_h4<CTA, false>(cs, G4, red);
  if (tid < G4) reinterpret_cast<float4*>(colsum)[(size_t)row * G4 + tid] = cs;
}

// SDDMM dS[e] = <g_i, z_j>. One warp per edge-slice: light rows = one warp per row; hub rows = CTA per row, warps split the edges.
template <bool CTA>
__global__ void sddmm_kernel(const RowSel r, const uint32_t* __restrict__ colidx, int F, size_t ld, const float* __restrict__ grad, const float* __restrict__ z,
                             float* __restrict__ ds, int vec) {
  uint32_t row, s, e; int tid, nthr;
  if (!pick_row<CTA>(r, row, s, e, tid, nthr)) return;
  const int lane = threadIdx.x & 31;
  const int warp = CTA ? (threadIdx.x >> 5) : 0;
  const int nwarps = CTA ? (blockDim.x >> 5) : 1;
  const float* g = grad + (size_t)row * ld;
  if (vec == 4) {
    const int nch = F / 4;
    // keep up to 4 chunks of g_i in registers (F <= 512); further chunks are re-read (L1)
    float4 gr[4];
#pragma unroll
    for (int k = 0; k < 4; k++) gr[k] = (lane + 32 * k < nch) ? __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * k) : make_float4(0, 0, 0, 0);
    for (uint32_t k0 = s + warp * 2; k0 < e; k0 += nwarps * 2) {
      float d[2] = {0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (k0 + u < e) {
          const float4* x = reinterpret_cast<const float4*>(z + (size_t)__ldg(colidx + k0 + u) * ld);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            if (lane + 32 * k < nch) { const float4 v = __ldg(x + lane + 32 * k); d[u] += gr[k].x * v.x + gr[k].y * v.y + gr[k].z * v.z + gr[k].w * v.w; }
          }
          for (int c = lane + 128; c < nch; c += 32) { const float4 v = __ldg(x + c); const float4 q = __ldg(reinterpret_cast<const float4*>(g) + c); d[u] += q.x * v.x + q.y * v.y + q.z * v.z + q.w * v.w; }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const float t = warp_sum(d[u]);
        if (lane == 0 && k0 + u < e) ds[k0 + u] = t;
      }
    }
  } else {
    for (uint32_t k0 = s + warp; k0 < e; k0 += nwarps) {
      const float* x = z + (size_t)__ldg(colidx + k0) * ld;
      float d = 0.f;
      for (int c = lane; c < F; c += 32) d += __ldg(g + c) * __ldg(x + c);
      d = warp_sum(d);
      if (lane == 0) ds[k0] = d;
    }
  }
}

// Edge-parallel SDDMM: dS[e] = <g_row(e), z_col(e)> over FLAT edge ranges — the products are independent, so unlike the aggregation
// there is no per-row order to respect and the work is cut into equal chunks of 1024 edges (perfect balance on power-law graphs, no
// hub / light split). A warp walks its chunk in batches of 32 edges: the neighbour rows of U edges are requested first (they do not
// depend on the row the edge belongs to), then each edge's partial dot product is formed against the row's gradient chunks held in
// registers (reloaded when the edge index crosses a row boundary, a warp-uniform test), and the 32 x 32 partials are reduced with a
// butterfly reduce-scatter: 31 shuffles per 32 edges instead of 5 per edge, results stored coalesced.
template <int KCH>
__global__ void __launch_bounds__(256, 4) sddmm_edges_kernel(uint32_t nv, uint64_t nnz, const uint32_t* __restrict__ rowptr,
                                                             const uint32_t* __restrict__ colidx, int nch, size_t ld4,
                                                             const float4* __restrict__ grad4, const float4* __restrict__ z4, float* __restrict__ ds) {
  constexpr int EC = 1024;
  constexpr int EB = 16;   // edges per batch (one reduce-scatter): 16 keeps the kernel at 64 registers = 4 CTAs per SM
  constexpr int U = KCH == 1 ? 8 : (KCH == 2 ? 4 : 2);
  const int lane = threadIdx.x & 31;
  const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t nchunks = (nnz + EC - 1) / EC;
  bool act[KCH];
#pragma unroll
  for (int k = 0; k < KCH; k++) act[k] = lane + 32 * k < nch;
  for (uint64_t chunk = gwarp; chunk < nchunks; chunk += nwarps) {
    const uint64_t e0 = chunk * EC;
    const uint64_t e1 = e0 + EC < nnz ? e0 + EC : nnz;
    // row containing edge e0: the last row with rowptr[row] <= e0 (binary search, warp-uniform)
    uint32_t lo = 0, hi = nv;
    while (hi - lo > 1) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if ((uint64_t)__ldg(rowptr + mid) <= e0) lo = mid; else hi = mid;
    }
    uint32_t row = lo;
    uint64_t re = __ldg(rowptr + row + 1);
    while (re <= e0) { row++; re = __ldg(rowptr + row + 1); }  // empty rows share their start with the next one
    float4 gq[KCH];
#pragma unroll
    for (int k = 0; k < KCH; k++) gq[k] = act[k] ? __ldg(grad4 + (size_t)row * ld4 + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint64_t b = e0; b < e1; b += EB) {
      const uint32_t c = (lane < EB && b + lane < e1) ? __ldg(colidx + b + lane) : 0u;
      float p[EB];
#pragma unroll
      for (int j0 = 0; j0 < EB; j0 += U) {
        float4 x[U][KCH];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t cc = __shfl_sync(0xffffffffu, c, j0 + u);
#pragma unroll
          for (int k = 0; k < KCH; k++) x[u][k] = (act[k] && b + j0 + u < e1) ? __ldg(z4 + (size_t)cc * ld4 + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint64_t edge = b + j0 + u;
          if (edge < e1 && edge >= re) {  // warp-uniform: the edge starts a new row
            do { row++; re = __ldg(rowptr + row + 1); } while (edge >= re);
#pragma unroll
            for (int k = 0; k < KCH; k++) gq[k] = act[k] ? __ldg(grad4 + (size_t)row * ld4 + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float d = 0.f;
#pragma unroll
          for (int k = 0; k < KCH; k++) d += gq[k].x * x[u][k].x + gq[k].y * x[u][k].y + gq[k].z * x[u][k].z + gq[k].w * x[u][k].w;
          p[j0 + u] = d;
        }
      }
      // butterfly reduce-scatter: lane L ends with the total of edge b + (L mod EB) in p[0]
#pragma unroll
      for (int o = EB / 2; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; i++) {
          const float send = up ? p[i] : p[i + o];
          const float keep = up ? p[i + o] : p[i];
          p[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      if (EB < 32) p[0] += __shfl_xor_sync(0xffffffffu, p[0], 16);  // lanes L and L + 16 hold the two halves of edge b + (L & 15)
      if (lane < EB && b + lane < e1) ds[b + lane] = p[0];
    }
  }
}

// Multi-head SDDMM: dS[e * H + h] = <g_row(e)[head h], z_col(e)[head h]>. Same flat edge chunks as above; a head owns `cph` consecutive
// float4 chunks (a power of two <= 32), i.e. `cph` consecutive lanes of a 32-chunk slab, so its dot product is a segmented butterfly
// (log2(cph) shuffles) and the 32 / cph heads of a slab are stored by their first lanes as one contiguous run.
template <int KCH>
__global__ void __launch_bounds__(256, 4) sddmm_heads_kernel(uint32_t nv, uint64_t nnz, const uint32_t* __restrict__ rowptr,
                                                             const uint32_t* __restrict__ colidx, int nch, int cph, int H, size_t ld4,
                                                             const float4* __restrict__ grad4, const float4* __restrict__ z4, float* __restrict__ ds) {
  constexpr int EC = 1024;
  constexpr int U = KCH == 1 ? 8 : (KCH == 2 ? 4 : 2);
  const int lane = threadIdx.x & 31;
  const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t nchunks = (nnz + EC - 1) / EC;
  bool act[KCH];
  int head[KCH];
#pragma unroll
  for (int k = 0; k < KCH; k++) { act[k] = lane + 32 * k < nch; head[k] = (lane + 32 * k) / cph; }
  const bool writer = (lane & (cph - 1)) == 0;
  for (uint64_t chunk = gwarp; chunk < nchunks; chunk += nwarps) {
    const uint64_t e0 = chunk * EC;
    const uint64_t e1 = e0 + EC < nnz ? e0 + EC : nnz;
    uint32_t lo = 0, hi = nv;
    while (hi - lo > 1) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if ((uint64_t)__ldg(rowptr + mid) <= e0) lo = mid; else hi = mid;
    }
    uint32_t row = lo;
    uint64_t re = __ldg(rowptr + row + 1);
    while (re <= e0) { row++; re = __ldg(rowptr + row + 1); }
    float4 gq[KCH];
#pragma unroll
    for (int k = 0; k < KCH; k++) gq[k] = act[k] ? __ldg(grad4 + (size_t)row * ld4 + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint64_t b = e0; b < e1; b += 32) {
      const uint32_t c = (b + lane < e1) ? __ldg(colidx + b + lane) : 0u;
#pragma unroll 1
      for (int j0 = 0; j0 < 32; j0 += U) {
        if (b + j0 >= e1) break;
        float4 x[U][KCH];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t cc = __shfl_sync(0xffffffffu, c, j0 + u);
#pragma unroll
          for (int k = 0; k < KCH; k++) x[u][k] = (act[k] && b + j0 + u < e1) ? __ldg(z4 + (size_t)cc * ld4 + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float d[U][KCH];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint64_t edge = b + j0 + u;
          if (edge < e1 && edge >= re) {  // warp-uniform: the edge starts a new row
            do { row++; re = __ldg(rowptr + row + 1); } while (edge >= re);
#pragma unroll
            for (int k = 0; k < KCH; k++) gq[k] = act[k] ? __ldg(grad4 + (size_t)row * ld4 + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int k = 0; k < KCH; k++) d[u][k] = gq[k].x * x[u][k].x + gq[k].y * x[u][k].y + gq[k].z * x[u][k].z + gq[k].w * x[u][k].w;
        }
        // segmented butterflies of the U x KCH partial sums, round by round (independent shuffles in flight instead of one chain per value)
        for (int o = cph >> 1; o >= 1; o >>= 1) {
#pragma unroll
          for (int u = 0; u < U; u++)
#pragma unroll
            for (int k = 0; k < KCH; k++) d[u][k] += __shfl_xor_sync(0xffffffffu, d[u][k], o);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint64_t edge = b + j0 + u;
#pragma unroll
          for (int k = 0; k < KCH; k++)
            if (writer && act[k] && edge < e1) ds[edge * H + head[k]] = d[u][k];
        }
      }
    }
  }
}

// In place on ds: softmax backward (closed form of math_functions.cpp:496-514), LeakyReLU backward
// (gat_aggregator.cpp:144); rowsum[i] = sum_e ds_e (the reference's src_score_grad, :149).
template <bool CTA>
__global__ void softmax_bwd_kernel(const RowSel r, int H, float slope, const float* __restrict__ temp_scores, const float* __restrict__ p,
                                   float* __restrict__ ds, float* __restrict__ rowsum) {
  __shared__ float red[32 * 8];
  uint32_t row, s, e; int tid, nthr;
  if (!pick_row<CTA>(r, row, s, e, tid, nthr)) return;
  const uint64_t i0 = (uint64_t)s * H, i1 = (uint64_t)e * H;
  float dot = 0.f;
  for (uint64_t i = i0 + tid; i < i1; i += nthr) dot += p[i] * ds[i];
  dot = coop_hsum<CTA>(dot, H, red);
  float rs = 0.f;
  for (uint64_t i = i0 + tid; i < i1; i += nthr) {
    const float dy = p[i] * (ds[i] - dot);
    const float v = dy * (temp_scores[i] > 0.f ? 1.0f : slope);
    ds[i] = v;
    rs += v;
  }
  rs = coop_hsum<CTA>(rs, H, red);
  if (tid < H) rowsum[(size_t)row * H + tid] = rs;
}

// colsum[j] = sum over edges e' pointing at j of ds[e'] = sum_{e in row j} ds[perm[e]]  (symmetric pattern)
template <bool CTA>
__global__ void colsum_kernel(const RowSel r, int H, const uint32_t* __restrict__ perm, const float* __restrict__ ds, float* __restrict__ colsum) {
  __shared__ float red[32 * 8];
  uint32_t row, s, e; int tid, nthr;
  if (!pick_row<CTA>(r, row, s, e, tid, nthr)) return;
  const int h = tid & (H - 1), hs = __ffs(H) - 1;
  const uint64_t i0 = (uint64_t)s * H, i1 = (uint64_t)e * H;
  float cs = 0.f;
  for (uint64_t i = i0 + tid; i < i1; i += nthr) cs += __ldg(ds + (size_t)__ldg(perm + (i >> hs)) * H + h);
  cs = coop_hsum<CTA>(cs, H, red);
  if (tid < H) colsum[(size_t)row * H + tid] = cs;
}

// Stage 1 of d_alpha = Z^T·[rowsum colsum]: each CTA reduces a slab of rows; thread (tx, ty): column tx (+CW*q), rows ty, ty+RH, ...
__global__ void alpha_grad_stage1(uint32_t nv, int F, int H, size_t ld, const float* __restrict__ z, const float* __restrict__ rowsum, const float* __restrict__ colsum,
                                  uint32_t rows_per_cta, int CW, float* __restrict__ partial /*[grid][2][F]*/) {
  extern __shared__ float sm[];  // [RH][2][CW]
  const int tx = threadIdx.x % CW, ty = threadIdx.x / CW, RH = blockDim.x / CW;
  const uint32_t r0 = blockIdx.x * rows_per_cta;
  const uint32_t r1 = (r0 + rows_per_cta < nv) ? r0 + rows_per_cta : nv;
  for (int c0 = 0; c0 < F; c0 += CW) {
    const int c = c0 + tx;
    float al = 0.f, ar = 0.f;
    if (c < F) {
      const int h = c / (F / H);   // the head this column belongs to
      for (uint32_t i = r0 + ty; i < r1; i += RH) {
        const float v = __ldg(z + (size_t)i * ld + c);
        al += __ldg(rowsum + (size_t)i * H + h) * v;
        ar += __ldg(colsum + (size_t)i * H + h) * v;
      }
    }
    sm[(ty * 2 + 0) * CW + tx] = al;
    sm[(ty * 2 + 1) * CW + tx] = ar;
    __syncthreads();
    if (ty == 0 && c < F) {
      float sl = 0.f, sr = 0.f;
      for (int q = 0; q < RH; q++) { sl += sm[(q * 2 + 0) * CW + tx]; sr += sm[(q * 2 + 1) * CW + tx]; }
      partial[((size_t)blockIdx.x * 2 + 0) * F + c] = sl;
      partial[((size_t)blockIdx.x * 2 + 1) * F + c] = sr;
    }
    __syncthreads();
  }
}
__global__ void alpha_grad_stage2(int F, int nparts, const float* __restrict__ partial, float* __restrict__ dal, float* __restrict__ dar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * F) return;
  const int which = i / F, c = i % F;
  float s = 0.f;
  for (int p = 0; p < nparts; p++) s += partial[((size_t)p * 2 + which) * F + c];
  (which ? dar : dal)[c] = s;
}

RowSel make_sel(gai_csr_t g) {
  RowSel r;
  r.rowptr = g->rowptr; r.hub_rows = g->hub_rows; r.order = g->row_order; r.nv = g->nv; r.n_hub = g->n_hub;
  r.hub_threshold = g->n_hub ? g->hub_degree : 0xffffffffu;
  return r;
}
inline unsigned warp_grid(uint32_t nv) { return (unsigned)(((uint64_t)nv * 32 + 255) / 256); }

}  // namespace

extern "C" {

static bool heads_ok(int F, int H) {
  // a power-of-two number of heads <= 32; for H > 1 whole float4 chunks per head, a power-of-two number (<= 32) of them
  if (H < 1 || H > 32 || (H & (H - 1)) != 0 || F % H != 0) return false;
  if (H == 1) return true;
  const int D = F / H;
  return D % 4 == 0 && ((D / 4) & (D / 4 - 1)) == 0 && D / 4 <= 32 && F <= 512;
}

int gai_gat_forward_heads_ld(gai_csr_t g, int F, int H, const float* z, size_t ld, const float* alpha_l, const float* alpha_r, float slope,
                             float* temp_scores, float* norm_scores, float* out, size_t ld_out, int flags, gai_stream_t stream) {
  GAI_CHECK_ARG(g && z && alpha_l && alpha_r && temp_scores && norm_scores && out && F > 0 && ld >= (size_t)F && ld_out >= (size_t)F);
  GAI_CHECK_ARG(heads_ok(F, H));
  if (g->nv == 0) return GAI_OK;
  cudaStream_t st = gai::S(stream);
  void* ws = nullptr;
  int rc = gai::workspace(sizeof(float) * 2 * (size_t)g->nv * H, &ws, st);
  if (rc != GAI_OK) return rc;
  float* el = reinterpret_cast<float*>(ws);
  float* er = el + (size_t)g->nv * H;
  el_er_kernel<<<warp_grid(g->nv), 256, 0, st>>>(g->nv, F, H, z, ld, alpha_l, alpha_r, el, er);
  GAI_LAUNCH_CHECK();
  const RowSel r = make_sel(g);
  const bool v4 = H % 4 == 0 && reinterpret_cast<uintptr_t>(temp_scores) % 16 == 0 && reinterpret_cast<uintptr_t>(norm_scores) % 16 == 0;
  if (v4) scores_kernel_v4<false><<<warp_grid(g->nv), 256, 0, st>>>(r, H, g->colidx, el, er, slope, temp_scores, norm_scores);
  else scores_kernel<false><<<warp_grid(g->nv), 256, 0, st>>>(r, H, g->colidx, el, er, slope, temp_scores, norm_scores);
  GAI_LAUNCH_CHECK();
  if (g->n_hub) {
    if (v4) scores_kernel_v4<true><<<g->n_hub, 256, 0, st>>>(r, H, g->colidx, el, er, slope, temp_scores, norm_scores);
    else scores_kernel<true><<<g->n_hub, 256, 0, st>>>(r, H, g->colidx, el, er, slope, temp_scores, norm_scores);
    GAI_LAUNCH_CHECK();
  }
  if (H == 1) return gai_spmm_edge(g, F, norm_scores, nullptr, z, (int)ld, out, (int)ld_out, flags, nullptr, stream);
  return gai_spmm_edge_heads(g, F, H, norm_scores, nullptr, z, (int)ld, out, (int)ld_out, flags, nullptr, stream);
}

int gai_gat_forward_ld(gai_csr_t g, int F, const float* z, size_t ld, const float* alpha_l, const float* alpha_r, float slope, float* temp_scores,
                       float* norm_scores, float* out, size_t ld_out, int flags, gai_stream_t stream) {
  return gai_gat_forward_heads_ld(g, F, 1, z, ld, alpha_l, alpha_r, slope, temp_scores, norm_scores, out, ld_out, flags, stream);
}

int gai_gat_backward_ld(gai_csr_t g, int F, const float* z, size_t ld, const float* grad_in, size_t ld_grad, float slope, const float* temp_scores,
                        const float* norm_scores, float* ds, float* d_alpha_l, float* d_alpha_r, float* dz, size_t ld_dz, gai_stream_t stream) {
  return gai_gat_backward_heads_ld(g, F, 1, z, ld, grad_in, ld_grad, slope, temp_scores, norm_scores, ds, d_alpha_l, d_alpha_r, dz, ld_dz, stream);
}

int gai_gat_backward_heads_ld(gai_csr_t g, int F, int H, const float* z, size_t ld, const float* grad_in, size_t ld_grad, float slope,
                              const float* temp_scores, const float* norm_scores, float* ds, float* d_alpha_l, float* d_alpha_r, float* dz, size_t ld_dz,
                              gai_stream_t stream) {
  GAI_CHECK_ARG(g && z && grad_in && temp_scores && norm_scores && ds && d_alpha_l && d_alpha_r && dz && F > 0);
  GAI_CHECK_ARG(ld >= (size_t)F && ld_grad >= (size_t)F && ld_dz >= (size_t)F);
  GAI_CHECK_ARG(heads_ok(F, H));
  if (g->nv == 0) return GAI_OK;
  int rc = gai_csr_build_transpose(g, stream);
  if (rc != GAI_OK) return rc;
  cudaStream_t st = gai::S(stream);
  const int sms = gai::sm_count();
  const int nparts = (int)((g->nv + 255) / 256 < (uint32_t)(4 * sms) ? (g->nv + 255) / 256 : (uint32_t)(4 * sms));
  void* ws = nullptr;
  rc = gai::workspace(sizeof(float) * (2 * (size_t)g->nv * H + (size_t)nparts * 2 * F), &ws, st);
  if (rc != GAI_OK) return rc;
  float* rowsum = reinterpret_cast<float*>(ws);
  float* colsum = rowsum + (size_t)g->nv * H;
  float* partial = colsum + (size_t)g->nv * H;
  const RowSel r = make_sel(g);
  // 128-bit path: both gathered matrices share one pitch that is a multiple of 4 floats (the tail chunk of a width that is not reads
  // padding columns, which the layer classes keep at zero on both sides: 0 * 0 adds nothing to the dot product)
  const bool vec_ok = ld == ld_grad && ld % 4 == 0 && ld >= (size_t)((F + 3) / 4 * 4) && reinterpret_cast<uintptr_t>(z) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(grad_in) % 16 == 0;
  if (H > 1) {
    // multi-head: both gathered matrices must be 128-bit loadable with one pitch (the layer classes' pitched buffers are)
    if (!vec_ok) return gai::set_error(GAI_ERR_ARG, "gai_gat_backward_heads", "H > 1 needs 16-byte aligned rows with one pitch that is a multiple of 4 floats");
    if (g->nnz > 0) {
      const int nch = F / 4, cph = F / H / 4;
      const unsigned grid = (unsigned)(sms * 4);
      const float4* g4 = reinterpret_cast<const float4*>(grad_in);
      const float4* z4 = reinterpret_cast<const float4*>(z);
      if (nch <= 32) sddmm_heads_kernel<1><<<grid, 256, 0, st>>>(g->nv, g->nnz, g->rowptr, g->colidx, nch, cph, H, ld / 4, g4, z4, ds);
      else if (nch <= 64) sddmm_heads_kernel<2><<<grid, 256, 0, st>>>(g->nv, g->nnz, g->rowptr, g->colidx, nch, cph, H, ld / 4, g4, z4, ds);
      else sddmm_heads_kernel<4><<<grid, 256, 0, st>>>(g->nv, g->nnz, g->rowptr, g->colidx, nch, cph, H, ld / 4, g4, z4, ds);
      GAI_LAUNCH_CHECK();
    }
  } else if (vec_ok && F <= 512 && g->nnz > 0) {
    const int nch = (F + 3) / 4;
    const unsigned grid = (unsigned)(sms * 4);
    const float4* g4 = reinterpret_cast<const float4*>(grad_in);
    const float4* z4 = reinterpret_cast<const float4*>(z);
    if (nch <= 32) sddmm_edges_kernel<1><<<grid, 256, 0, st>>>(g->nv, g->nnz, g->rowptr, g->colidx, nch, ld / 4, g4, z4, ds);
    else if (nch <= 64) sddmm_edges_kernel<2><<<grid, 256, 0, st>>>(g->nv, g->nnz, g->rowptr, g->colidx, nch, ld / 4, g4, z4, ds);
    else sddmm_edges_kernel<4><<<grid, 256, 0, st>>>(g->nv, g->nnz, g->rowptr, g->colidx, nch, ld / 4, g4, z4, ds);
    GAI_LAUNCH_CHECK();
  } else {
    const int vec = (vec_ok && F % 4 == 0) ? 4 : 1;
    const size_t ldc = ld;  // the fallback walks both matrices with the pitch of z; grad_in must share it
    GAI_CHECK_ARG(ld == ld_grad);
    sddmm_kernel<false><<<warp_grid(g->nv), 256, 0, st>>>(r, g->colidx, F, ldc, grad_in, z, ds, vec);
    GAI_LAUNCH_CHECK();
    if (g->n_hub) { sddmm_kernel<true><<<g->n_hub, 256, 0, st>>>(r, g->colidx, F, ldc, grad_in, z, ds, vec); GAI_LAUNCH_CHECK(); }
  }
  const bool v4 = H % 4 == 0 && reinterpret_cast<uintptr_t>(temp_scores) % 16 == 0 && reinterpret_cast<uintptr_t>(norm_scores) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(ds) % 16 == 0;
  if (v4) softmax_bwd_kernel_v4<false><<<warp_grid(g->nv), 256, 0, st>>>(r, H, slope, temp_scores, norm_scores, ds, rowsum);
  else softmax_bwd_kernel<false><<<warp_grid(g->nv), 256, 0, st>>>(r, H, slope, temp_scores, norm_scores, ds, rowsum);
  GAI_LAUNCH_CHECK();
  if (g->n_hub) {
    if (v4) softmax_bwd_kernel_v4<true><<<g->n_hub, 256, 0, st>>>(r, H, slope, temp_scores, norm_scores, ds, rowsum);
    else softmax_bwd_kernel<true><<<g->n_hub, 256, 0, st>>>(r, H, slope, temp_scores, norm_scores, ds, rowsum);
    GAI_LAUNCH_CHECK();
  }
  if (v4) colsum_kernel_v4<false><<<warp_grid(g->nv), 256, 0, st>>>(r, H, g->tperm, ds, colsum);
  else colsum_kernel<false><<<warp_grid(g->nv), 256, 0, st>>>(r, H, g->tperm, ds, colsum);
  GAI_LAUNCH_CHECK();
  if (g->n_hub) {
    if (v4) colsum_kernel_v4<true><<<g->n_hub, 256, 0, st>>>(r, H, g->tperm, ds, colsum);
    else colsum_kernel<true><<<g->n_hub, 256, 0, st>>>(r, H, g->tperm, ds, colsum);
    GAI_LAUNCH_CHECK();
  }
  int CW = 1;
  while (CW < F && CW < 256) CW <<= 1;
  const uint32_t rows_per_cta = (g->nv + nparts - 1) / nparts;
  alpha_grad_stage1<<<nparts, 256, sizeof(float) * 2 * 256, st>>>(g->nv, F, H, ld, z, rowsum, colsum, rows_per_cta, CW, partial);
  GAI_LAUNCH_CHECK();
  alpha_grad_stage2<<<(2 * F + 255) / 256, 256, 0, st>>>(F, nparts, partial, d_alpha_l, d_alpha_r);
  GAI_LAUNCH_CHECK();
  // dZ = P^T · G  (update_all with transposed scores, gat_aggregator.cpp:175-199); z is dead from here on, dz may alias it
  if (H == 1) return gai_spmm_edge(g, F, norm_scores, g->tperm, grad_in, (int)ld_grad, dz, (int)ld_dz, GAI_EPI_NONE, nullptr, stream);
  return gai_spmm_edge_heads(g, F, H, norm_scores, g->tperm, grad_in, (int)ld_grad, dz, (int)ld_dz, GAI_EPI_NONE, nullptr, stream);
}

int gai_gat_forward(gai_csr_t g, int F, const float* z, const float* alpha_l, const float* alpha_r, float slope, float* temp_scores,
                    float* norm_scores, float* out, int flags, gai_stream_t stream) {
  return gai_gat_forward_ld(g, F, z, (size_t)F, alpha_l, alpha_r, slope, temp_scores, norm_scores, out, (size_t)F, flags, stream);
}

int gai_gat_backward(gai_csr_t g, int F, const float* z, const float* grad_in, float slope, const float* temp_scores, const float* norm_scores,
                     float* ds, float* d_alpha_l, float* d_alpha_r, float* dz, gai_stream_t stream) {
  return gai_gat_backward_ld(g, F, z, (size_t)F, grad_in, (size_t)F, slope, temp_scores, norm_scores, ds, d_alpha_l, d_alpha_r, dz, (size_t)F, stream);
}

}  // extern "C"
