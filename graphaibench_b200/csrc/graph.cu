// Device CSR for the GNN path: upload, degree normalisers, hub-row list, transpose permutation.
// Replaces LearningGraph::alloc_on_device/copy_to_gpu/compute_vertex_data (src/gnn/lgraph.cu:51-105) and the
// per-backward cusparseCsr2cscEx2 call (src/utilities/math_functions.cu:345-358, src/gnn/gconv/gat_aggregator.cu:86-89).
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "gai_internal.cuh"

namespace {

// One thread per vertex. Both normalisers reproduce the reference's mixed float/double expressions exactly:
//   lgraph.cpp:29-32        temp = sqrtf(float(deg)); v = temp == 0 ? 0 : float(1.0 / double(temp))
//   sage_aggregator.cpp:17  b = float(1.0 / double(float(deg)))        (inf for deg == 0, never used by a row of its own)
__global__ void norms_kernel(uint32_t nv, const uint32_t* __restrict__ rowptr, float* __restrict__ norm_gcn, float* __restrict__ norm_mean) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  uint32_t deg = rowptr[v + 1] - rowptr[v];
  float fdeg = __uint2float_rn(deg);
  float t = __fsqrt_rn(fdeg);
  norm_gcn[v] = (t == 0.0f) ? 0.0f : __double2float_rn(__ddiv_rn(1.0, (double)t));
  norm_mean[v] = __double2float_rn(__ddiv_rn(1.0, (double)fdeg));
}

// One warp per row; each lane binary-searches its edges' mirror position. Pattern must be symmetric
// with sorted rows (what the reference's symmetric_csr_transpose assumes, math_functions.cpp:32-74).
__global__ void transpose_perm_kernel(uint32_t nv, const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
                                      uint32_t* __restrict__ perm, uint32_t* __restrict__ bad) {
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  if (warp >= nv) return;
  uint32_t src = warp;
  uint32_t b = rowptr[src], e = rowptr[src + 1];
  for (uint32_t k = b + lane; k < e; k += 32) {
    uint32_t dst = colidx[k];
    int64_t l = rowptr[dst], r = (int64_t)rowptr[dst + 1] - 1;
    int64_t found = -1;
    while (r >= l) {
      int64_t mid = l + (r - l) / 2;
      uint32_t val = colidx[mid];
      if (val == src) { found = mid; break; }
      if (val < src) l = mid + 1; else r = mid - 1;
    }
    if (found < 0) { atomicAdd(bad, 1u); perm[k] = (uint32_t)k; }
    else perm[k] = (uint32_t)found;
  }
}

// Work lists for the aggregation kernels (integer work on the host, once per graph):
//   row_order  all rows, longest first, ascending id among equal degrees (counting sort, O(nv + max degree));
//              its first n_hub entries (degree > hub_degree) are the hub rows, served one CTA each;
//   claim_ptr  the remaining (light) rows cut into claims of at most 32 rows and at most CLAIM_EDGES edges, stored as
//              (begin, end) pairs in execution order: the unit a warp takes from the shared work counter.
constexpr uint64_t CLAIM_EDGES = 2048;

// Builds the list of rows [rb, re) into `out` (device arrays owned by `out`).
int build_list(const gai_csr* g, const uint32_t* rowptr_h, uint32_t rb, uint32_t re, gai_worklist* out, cudaStream_t st) {
  out->rb = rb; out->re = re; out->n_hub = 0; out->n_claims = 0;
  const uint32_t n = re - rb;
  if (n == 0) return GAI_OK;
  auto deg = [&](uint32_t v) { return rowptr_h[v + 1] - rowptr_h[v]; };
  uint32_t maxdeg = 0;
  for (uint32_t v = rb; v < re; v++) maxdeg = std::max(maxdeg, deg(v));
  std::vector<uint64_t> start((size_t)maxdeg + 2, 0);
  for (uint32_t v = rb; v < re; v++) start[(size_t)maxdeg - deg(v) + 1]++;  // bucket 0 = longest
  for (size_t d = 1; d < start.size(); d++) start[d] += start[d - 1];
  std::vector<uint32_t> order(n);
  for (uint32_t v = rb; v < re; v++) order[start[(size_t)maxdeg - deg(v)]++] = v;
  uint32_t n_hub = 0;
  while (n_hub < n && deg(order[n_hub]) > g->hub_degree) n_hub++;
  // claims as (begin, end) pairs into the light part of the list, in EXECUTION order
  std::vector<uint32_t> cuts;
  cuts.reserve((n - n_hub) / 16 + 2);
  cuts.push_back(0);
  uint64_t edges = 0;
  uint32_t rows = 0, n_big = 0;
  for (uint32_t i = n_hub; i < n; i++) {
    const uint32_t d = deg(order[i]);
    if (rows > 0 && (rows == 32 || edges + d > CLAIM_EDGES)) { cuts.push_back(i - n_hub); rows = 0; edges = 0; }
    if (d > CLAIM_EDGES) n_big++;  // rows longer than the budget travel alone
    rows++; edges += d;
  }
  cuts.push_back(n - n_hub);
  const uint32_t n_claims = (uint32_t)cuts.size() - 1;
  // Execution order: the over-budget single-row claims first (longest-processing-time first keeps the tail short), then
  // the rest interleaved by a stride permutation so that, at any moment, the machine works on a mix of long-row claims
  // (bandwidth-bound) and short-row claims (latency-bound) instead of one degree class at a time.
  std::vector<uint32_t> claims((size_t)2 * n_claims);
  const char* env = getenv("GAI_CLAIM_ORDER");
  const int mode = env ? atoi(env) : 1;
  const uint32_t rest = n_claims - n_big;
  uint64_t stride = 1;
  if (mode == 1 && rest > 64) {
    static const uint64_t primes[] = {1000003ull, 998244353ull, 2654435761ull, 40503ull, 7919ull};
    for (uint64_t pr : primes) {
      uint64_t x = pr % rest, y = rest;
      while (y) { const uint64_t t = x % y; x = y; y = t; }
      if (x == 1 && (pr % rest) > 1) { stride = pr % rest; break; }
    }
  }
  for (uint32_t i = 0; i < n_claims; i++) {
    const uint32_t src = i < n_big ? i : n_big + (uint32_t)(((uint64_t)(i - n_big) * stride) % rest);
    claims[2 * (size_t)i] = cuts[src];
    claims[2 * (size_t)i + 1] = cuts[src + 1];
  }
  out->n_hub = n_hub;
  out->n_claims = n_claims;
  GAI_CUDA(cudaMalloc(&out->row_order, sizeof(uint32_t) * (size_t)n));
  GAI_CUDA(cudaMalloc(&out->claim_ptr, sizeof(uint32_t) * claims.size()));
  GAI_CUDA(cudaMemcpyAsync(out->row_order, order.data(), sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  GAI_CUDA(cudaMemcpyAsync(out->claim_ptr, claims.data(), sizeof(uint32_t) * claims.size(), cudaMemcpyHostToDevice, st));
  GAI_CUDA(cudaStreamSynchronize(st));  // the staging vectors are pageable and go out of scope
  return GAI_OK;
}

// rowptr_h == NULL: CSR supplied in device memory, bring the row pointers back once
int host_rowptr(const gai_csr* g, const uint32_t*& rowptr_h, std::vector<uint32_t>& keep, cudaStream_t st) {
  if (rowptr_h) return GAI_OK;
  keep.resize((size_t)g->nv + 1);
  GAI_CUDA(cudaMemcpyAsync(keep.data(), g->rowptr, sizeof(uint32_t) * ((size_t)g->nv + 1), cudaMemcpyDeviceToHost, st));
  GAI_CUDA(cudaStreamSynchronize(st));
  rowptr_h = keep.data();
  return GAI_OK;
}

int build_work_lists(gai_csr* g, const uint32_t* rowptr_h, cudaStream_t st) {
  g->n_hub = 0; g->n_claims = 0;
  if (g->nv == 0) return GAI_OK;
  std::vector<uint32_t> keep;
  int rc = host_rowptr(g, rowptr_h, keep, st);
  if (rc != GAI_OK) return rc;
  gai_worklist full;
  rc = build_list(g, rowptr_h, 0, g->nv, &full, st);
  if (rc != GAI_OK) return rc;
  g->row_order = full.row_order; g->claim_ptr = full.claim_ptr; g->n_hub = full.n_hub; g->n_claims = full.n_claims;
  for (uint32_t v = 0; v < g->nv; v++) g->max_degree = std::max(g->max_degree, rowptr_h[v + 1] - rowptr_h[v]);
  g->hub_rows = g->row_order;
  return GAI_OK;
}

int finish_create(gai_csr* g, cudaStream_t st, const uint32_t* rowptr_h) {
  GAI_CUDA(cudaMalloc(&g->norm_gcn, sizeof(float) * (size_t)(g->nv ? g->nv : 1)));
  GAI_CUDA(cudaMalloc(&g->norm_mean, sizeof(float) * (size_t)(g->nv ? g->nv : 1)));
  GAI_CUDA(cudaMalloc(&g->row_counters, sizeof(unsigned long long) * 16));
  g->hub_degree = gai::hub_degree_for(g->nnz);
  if (g->nv) {
    norms_kernel<<<(g->nv + 255) / 256, 256, 0, st>>>(g->nv, g->rowptr, g->norm_gcn, g->norm_mean);
    GAI_LAUNCH_CHECK();
  }
  return build_work_lists(g, rowptr_h, st);
}

}  // namespace

extern "C" {

int gai_add_selfloop_h(uint32_t nv, const uint32_t* rowptr, const uint32_t* colidx, uint32_t* rowptr_out, uint32_t* colidx_out) {
  GAI_CHECK_ARG(rowptr && rowptr_out && (colidx || rowptr[nv] == 0) && colidx_out);
  // Row i of the output is row i of the input with `i` placed in front of the first neighbour larger than i
  // (at the end if there is none) — the placement LearningGraph::add_selfloop makes (lgraph.h:185-218).
  for (uint32_t i = 0; i < nv; i++) {
    const uint32_t b = rowptr[i], e = rowptr[i + 1];
    uint32_t pos = e;
    for (uint32_t k = b; k < e; k++) {
      if (colidx[k] > i) { pos = k; break; }
    }
    uint32_t* o = colidx_out + (size_t)b + i;
    if (pos > b) memcpy(o, colidx + b, sizeof(uint32_t) * (pos - b));
    o[pos - b] = i;
    if (e > pos) memcpy(o + (pos - b) + 1, colidx + pos, sizeof(uint32_t) * (e - pos));
  }
  for (uint32_t i = nv + 1; i-- > 0;) rowptr_out[i] = rowptr[i] + i;  // descending: rowptr_out may alias rowptr
  return GAI_OK;
}

int gai_csr_create(uint32_t nv, uint64_t nnz, const uint32_t* rowptr_h, const uint32_t* colidx_h, gai_stream_t stream, gai_csr_t* out) {
  GAI_CHECK_ARG(out && rowptr_h && (colidx_h || nnz == 0));
  GAI_CHECK_ARG(nnz < (uint64_t(1) << 32));
  GAI_CHECK_ARG(rowptr_h[nv] == nnz);
  cudaStream_t st = gai::S(stream);
  gai_csr* g = new gai_csr();
  g->nv = nv; g->nnz = nnz; g->owns_csr = true;
  int rc = GAI_OK;
  do {
    if (cudaMalloc(&g->rowptr, sizeof(uint32_t) * ((size_t)nv + 1)) != cudaSuccess ||
        cudaMalloc(&g->colidx, sizeof(uint32_t) * (size_t)(nnz ? nnz : 1)) != cudaSuccess) {
      rc = gai::set_error(GAI_ERR_CUDA, "gai_csr_create", cudaGetErrorString(cudaGetLastError()));
      break;
    }
    if (cudaMemcpyAsync(g->rowptr, rowptr_h, sizeof(uint32_t) * ((size_t)nv + 1), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        (nnz && cudaMemcpyAsync(g->colidx, colidx_h, sizeof(uint32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st) != cudaSuccess)) {
      rc = gai::set_error(GAI_ERR_CUDA, "gai_csr_create upload", cudaGetErrorString(cudaGetLastError()));
      break;
    }
    rc = finish_create(g, st, rowptr_h);
  } while (0);
  if (rc != GAI_OK) { gai_csr_destroy(g); return rc; }
  *out = g;
  return GAI_OK;
}

int gai_csr_create_device(uint32_t nv, uint64_t nnz, const uint32_t* rowptr_d, const uint32_t* colidx_d, gai_stream_t stream, gai_csr_t* out) {
  GAI_CHECK_ARG(out && rowptr_d && (colidx_d || nnz == 0));
  GAI_CHECK_ARG(nnz < (uint64_t(1) << 32));
  gai_csr* g = new gai_csr();
  g->nv = nv; g->nnz = nnz; g->owns_csr = false;
  g->rowptr = const_cast<uint32_t*>(rowptr_d);
  g->colidx = const_cast<uint32_t*>(colidx_d);
  int rc = finish_create(g, gai::S(stream), nullptr);
  if (rc != GAI_OK) { gai_csr_destroy(g); return rc; }
  *out = g;
  return GAI_OK;
}

int gai_csr_destroy(gai_csr_t g) {
  if (!g) return GAI_OK;
  if (g->owns_csr) { cudaFree(g->rowptr); cudaFree(g->colidx); }
  for (int i = 0; i < g->n_seg; i++) { cudaFree(g->seg[i].row_order); cudaFree(g->seg[i].claim_ptr); }
  cudaFree(g->norm_gcn); cudaFree(g->norm_mean); cudaFree(g->tperm); cudaFree(g->row_counters); cudaFree(g->row_order); cudaFree(g->claim_ptr);
  if (g->aux_stream) { cudaStreamDestroy(g->aux_stream); cudaEventDestroy(g->ev_fork); cudaEventDestroy(g->ev_join); }
  delete g;
  return GAI_OK;
}

uint32_t gai_csr_nv(gai_csr_t g) { return g ? g->nv : 0; }
uint64_t gai_csr_nnz(gai_csr_t g) { return g ? g->nnz : 0; }
const uint32_t* gai_csr_rowptr(gai_csr_t g) { return g ? g->rowptr : nullptr; }
const uint32_t* gai_csr_colidx(gai_csr_t g) { return g ? g->colidx : nullptr; }
const float* gai_csr_vertex_norm(gai_csr_t g) { return g ? g->norm_gcn : nullptr; }
const float* gai_csr_mean_norm(gai_csr_t g) { return g ? g->norm_mean : nullptr; }
uint32_t gai_csr_num_hub_rows(gai_csr_t g) { return g ? g->n_hub : 0; }
uint32_t gai_csr_max_degree(gai_csr_t g) { return g ? g->max_degree : 0; }

int gai_csr_set_norms(gai_csr_t g, const float* norm_gcn_d, const float* norm_mean_d, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr);
  if (norm_gcn_d) GAI_CUDA(cudaMemcpyAsync(g->norm_gcn, norm_gcn_d, sizeof(float) * (size_t)g->nv, cudaMemcpyDeviceToDevice, gai::S(stream)));
  if (norm_mean_d) GAI_CUDA(cudaMemcpyAsync(g->norm_mean, norm_mean_d, sizeof(float) * (size_t)g->nv, cudaMemcpyDeviceToDevice, gai::S(stream)));
  return GAI_OK;
}

int gai_csr_set_row_segments(gai_csr_t g, int n_segments, const uint32_t* bounds_h, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr && n_segments >= 0 && n_segments <= GAI_MAX_SEGMENTS && (bounds_h || n_segments == 0));
  for (int i = 0; i < n_segments; i++) GAI_CHECK_ARG(bounds_h[2 * i] <= bounds_h[2 * i + 1] && bounds_h[2 * i + 1] <= g->nv);
  cudaStream_t st = gai::S(stream);
  for (int i = 0; i < g->n_seg; i++) { cudaFree(g->seg[i].row_order); cudaFree(g->seg[i].claim_ptr); g->seg[i] = gai_worklist(); }
  g->n_seg = 0;
  if (n_segments == 0 || g->nv == 0) return GAI_OK;
  const uint32_t* rowptr_h = nullptr;
  std::vector<uint32_t> keep;
  int rc = host_rowptr(g, rowptr_h, keep, st);
  if (rc != GAI_OK) return rc;
  for (int i = 0; i < n_segments; i++) {
    rc = build_list(g, rowptr_h, bounds_h[2 * i], bounds_h[2 * i + 1], &g->seg[i], st);
    if (rc != GAI_OK) return rc;
    g->n_seg = i + 1;
  }
  return GAI_OK;
}

int gai_csr_build_transpose(gai_csr_t g, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr);
  if (g->tperm) return GAI_OK;
  cudaStream_t st = gai::S(stream);
  GAI_CUDA(cudaMalloc(&g->tperm, sizeof(uint32_t) * (size_t)(g->nnz + 1)));
  uint32_t* bad = g->tperm + g->nnz;
  GAI_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), st));
  if (g->nv) {
    uint64_t threads = (uint64_t)g->nv * 32;
    transpose_perm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(g->nv, g->rowptr, g->colidx, g->tperm, bad);
    GAI_LAUNCH_CHECK();
  }
  uint32_t nbad = 0;
  GAI_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  GAI_CUDA(cudaStreamSynchronize(st));
  if (nbad) {
    cudaFree(g->tperm); g->tperm = nullptr;
    return gai::set_error(GAI_ERR_ARG, "gai_csr_build_transpose", "pattern is not structurally symmetric");
  }
  return GAI_OK;
}
const uint32_t* gai_csr_transpose_perm(gai_csr_t g) { return g ? g->tperm : nullptr; }

}  // extern "C"
