// CSR construction on the device: COO pairs -> sorted, de-duplicated CSR, and the self-loop insertion of the GNN path.
//
// Replaces the host set-based construction of the reference's text converter — Converter::read_mtx + adjlist2CSR
// (src/converters/converter.cc:27-60,314-420: one std::set of neighbours per vertex, self-loops dropped, the reverse edge added for
// symmetric inputs, rows emitted in ascending neighbour order, offsets by parallel_prefix_sum, include/scan.h:4-35) — and
// LearningGraph::add_selfloop (include/gnn/lgraph.h:185-218). Integer work, bit-exact: the result is the unique CSR of the edge SET,
// so any correct construction gives the reference's bytes; what changes is the schedule:
//   keys     one 64-bit key (src << 32 | dst) per kept pair (and its mirror if symmetrising); self-loops and out-of-range ids get the
//            all-ones sentinel, which sorts last
//   sort     LSD radix sort of the keys (cub::DeviceRadixSort — library code, like calling cuBLAS for a plain GEMM)
//   unique   cub::DeviceSelect::Unique on the sorted keys
//   offsets  rowptr[v] = first position whose key >= v << 32: one binary search per vertex (no degree histogram, no atomics, no scan)
//   columns  colidx[e] = low word of key e
// The same routine builds the benchmark graphs (graphaibench_b200/datagen.py) so that "CSR construction" is this code end to end.
#include <cub/cub.cuh>
#include "gai_internal.cuh"

namespace {

constexpr unsigned long long SENTINEL = ~0ull;

__global__ void make_keys_kernel(size_t n, const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst, uint32_t nv, int symmetrize,
                                 unsigned long long* __restrict__ keys) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t s = src[i], d = dst[i];
    const bool keep = s != d && s < nv && d < nv;  // converter.cc:372 "remove selfloops"; ids are asserted in range (converter.cc:52)
    keys[i] = keep ? ((unsigned long long)s << 32) | d : SENTINEL;
    if (symmetrize) keys[n + i] = keep ? ((unsigned long long)d << 32) | s : SENTINEL;
  }
}

// rowptr[v] = lower_bound(keys, v << 32) for v in [0, nv]; keys are sorted and unique (a trailing sentinel, if any, lies beyond n_keys)
__global__ void offsets_kernel(uint32_t nv, const unsigned long long* __restrict__ keys, size_t n_keys, int64_t* __restrict__ rowptr) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > nv) return;
  const unsigned long long want = (unsigned long long)v << 32;
  size_t lo = 0, hi = n_keys;
  while (lo < hi) {
    const size_t mid = lo + ((hi - lo) >> 1);
    if (keys[mid] < want) lo = mid + 1; else hi = mid;
  }
  rowptr[v] = (int64_t)lo;
}

__global__ void columns_kernel(size_t n, const unsigned long long* __restrict__ keys, uint32_t* __restrict__ colidx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) colidx[i] = (uint32_t)keys[i];
}

// add_selfloop (lgraph.h:185-218): row i gains `first + i` in front of its first neighbour larger than that id. One warp per row:
// the insertion point is the count of neighbours below the loop id (rows are sorted; no pre-existing loop, as the reference assumes).
__global__ void selfloop_kernel(uint32_t nv, uint32_t first, const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
                                uint32_t* __restrict__ rowptr_out, uint32_t* __restrict__ colidx_out) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w > nv) return;
  if (w == nv) { if (lane == 0) rowptr_out[nv] = rowptr[nv] + nv; return; }
  const uint32_t i = (uint32_t)w, self = first + i;
  const uint32_t b = rowptr[i], e = rowptr[i + 1];
  uint32_t lo = b, hi = e;  // first position whose neighbour is > self
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (colidx[mid] > self) hi = mid; else lo = mid + 1;
  }
  const uint32_t pos = lo;
  uint32_t* o = colidx_out + (size_t)b + i;
  for (uint32_t k = b + lane; k < e; k += 32) o[(k - b) + (k >= pos ? 1 : 0)] = colidx[k];
  if (lane == 0) { o[pos - b] = self; rowptr_out[i] = b + i; }
}

// ---- induced subgraph, re-indexed (Sampler::generateSubgraph: getMaskedGraph + reindexSubgraph, src/gnn/sampler.cpp:66-158) -----------
// new_id[v] = rank of v in the ascending keep list, 0xffffffff for vertices outside it
__global__ void scatter_ids_kernel(uint32_t n_keep, const uint32_t* __restrict__ keep, uint32_t* __restrict__ new_id) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_keep) new_id[keep[i]] = i;
}
// one warp per kept vertex. FILL = false: deg[i] = neighbours that are kept; FILL = true: their new ids, in the row's (ascending) order,
// written from out_rowptr[i] — a ballot prefix keeps the order inside each batch of 32 neighbours
template <bool FILL>
__global__ void induced_rows_kernel(uint32_t n_keep, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ rowptr,
                                    const uint32_t* __restrict__ colidx, const uint32_t* __restrict__ new_id, uint32_t* __restrict__ deg,
                                    const uint32_t* __restrict__ out_rowptr, uint32_t* __restrict__ out_colidx) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_keep) return;
  const uint32_t v = keep[w];
  const uint32_t b = rowptr[v], e = rowptr[v + 1];
  uint32_t count = 0;
  uint32_t* out = FILL ? out_colidx + out_rowptr[w] : nullptr;
  for (uint32_t k0 = b; k0 < e; k0 += 32) {
    const uint32_t k = k0 + lane;
    const uint32_t id = k < e ? new_id[colidx[k]] : 0xffffffffu;
    const unsigned kept = __ballot_sync(0xffffffffu, id != 0xffffffffu);
    if (FILL && id != 0xffffffffu) out[count + __popc(kept & ((1u << lane) - 1u))] = id;
    count += __popc(kept);
  }
  if (!FILL && lane == 0) deg[w] = count;
}

inline unsigned blocks_for(size_t n, int per_block = 256) {
  size_t b = (n + per_block - 1) / per_block;
  const size_t cap = (size_t)gai::sm_count() * 32;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int gai_coo_to_csr(uint32_t nv, uint64_t n_pairs, const uint32_t* src_d, const uint32_t* dst_d, int symmetrize, gai_stream_t stream,
                   int64_t** rowptr_out_d, uint32_t** colidx_out_d, uint64_t* nnz_out) {
  GAI_CHECK_ARG(rowptr_out_d != nullptr && colidx_out_d != nullptr && nnz_out != nullptr && (n_pairs == 0 || (src_d != nullptr && dst_d != nullptr)));
  cudaStream_t st = gai::S(stream);
  const size_t n_keys_in = (size_t)n_pairs * (symmetrize ? 2 : 1);
  *rowptr_out_d = nullptr; *colidx_out_d = nullptr; *nnz_out = 0;
  int64_t* rowptr = nullptr;
  GAI_CUDA(cudaMalloc(&rowptr, sizeof(int64_t) * ((size_t)nv + 1)));
  if (n_keys_in == 0) {
    GAI_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int64_t) * ((size_t)nv + 1), st));
    uint32_t* ci = nullptr;
    GAI_CUDA(cudaMalloc(&ci, sizeof(uint32_t)));
    *rowptr_out_d = rowptr; *colidx_out_d = ci;
    return GAI_OK;
  }
  unsigned long long *keys = nullptr, *keys_alt = nullptr, *d_num = nullptr;
  void* tmp = nullptr;
  int rc = GAI_OK;
  do {
    if (cudaMalloc(&keys, sizeof(unsigned long long) * n_keys_in) != cudaSuccess || cudaMalloc(&keys_alt, sizeof(unsigned long long) * n_keys_in) != cudaSuccess ||
        cudaMalloc(&d_num, sizeof(unsigned long long)) != cudaSuccess) {
      cudaGetLastError();
      rc = gai::set_error(GAI_ERR_NOMEM, "gai_coo_to_csr", "device memory for the sort keys");
      break;
    }
    make_keys_kernel<<<blocks_for(n_pairs), 256, 0, st>>>((size_t)n_pairs, src_d, dst_d, nv, symmetrize, keys);
    __atomic_fetch_add(&gai::g_launches, 1ull, __ATOMIC_RELAXED);
    cub::DoubleBuffer<unsigned long long> buf(keys, keys_alt);
    size_t tmp_bytes = 0, tmp_bytes2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, buf, (long long)n_keys_in, 0, 64, st);
    cub::DeviceSelect::Unique(nullptr, tmp_bytes2, keys, keys_alt, d_num, (long long)n_keys_in, st);
    if (tmp_bytes2 > tmp_bytes) tmp_bytes = tmp_bytes2;
    if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1) != cudaSuccess) { cudaGetLastError(); rc = gai::set_error(GAI_ERR_NOMEM, "gai_coo_to_csr", "sort workspace"); break; }
    // all 64 bits take part: the all-ones sentinel of dropped pairs then sorts behind every real key
    if (cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, buf, (long long)n_keys_in, 0, 64, st) != cudaSuccess) { rc = gai::set_error(GAI_ERR_CUDA, "gai_coo_to_csr", "radix sort"); break; }
    unsigned long long* sorted = buf.Current();
    unsigned long long* uniq = buf.Alternate();
    if (cub::DeviceSelect::Unique(tmp, tmp_bytes, sorted, uniq, d_num, (long long)n_keys_in, st) != cudaSuccess) { rc = gai::set_error(GAI_ERR_CUDA, "gai_coo_to_csr", "unique"); break; }
    unsigned long long n_unique = 0, last = 0;
    if (cudaMemcpyAsync(&n_unique, d_num, sizeof(n_unique), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = gai::set_error(GAI_ERR_CUDA, "gai_coo_to_csr", cudaGetErrorString(cudaGetLastError())); break; }
    if (n_unique) {
      if (cudaMemcpyAsync(&last, uniq + (n_unique - 1), sizeof(last), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = gai::set_error(GAI_ERR_CUDA, "gai_coo_to_csr", cudaGetErrorString(cudaGetLastError())); break; }
      if (last == SENTINEL) n_unique--;  // dropped pairs collapse into one trailing sentinel
    }
    if (n_unique >= (1ull << 32)) { rc = gai::set_error(GAI_ERR_UNSUPPORTED, "gai_coo_to_csr", "more than 2^32 - 1 edges"); break; }
    uint32_t* ci = nullptr;
    if (cudaMalloc(&ci, sizeof(uint32_t) * (n_unique ? n_unique : 1)) != cudaSuccess) { cudaGetLastError(); rc = gai::set_error(GAI_ERR_NOMEM, "gai_coo_to_csr", "column indices"); break; }
    offsets_kernel<<<(unsigned)(((size_t)nv + 1 + 255) / 256), 256, 0, st>>>(nv, uniq, (size_t)n_unique, rowptr);
    __atomic_fetch_add(&gai::g_launches, 1ull, __ATOMIC_RELAXED);
    if (n_unique) {
      columns_kernel<<<blocks_for((size_t)n_unique), 256, 0, st>>>((size_t)n_unique, uniq, ci);
      __atomic_fetch_add(&gai::g_launches, 1ull, __ATOMIC_RELAXED);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(ci); rc = gai::set_error(GAI_ERR_CUDA, "gai_coo_to_csr", cudaGetErrorString(cudaGetLastError())); break; }
    *rowptr_out_d = rowptr; *colidx_out_d = ci; *nnz_out = n_unique;
    rowptr = nullptr;
  } while (0);
  cudaFree(keys); cudaFree(keys_alt); cudaFree(d_num); cudaFree(tmp); cudaFree(rowptr);
  return rc;
}

int gai_induced_subgraph(gai_csr_t g, uint32_t n_keep, const uint32_t* keep_ids_d, gai_stream_t stream, uint32_t** rowptr_out_d, uint32_t** colidx_out_d,
                         uint64_t* nnz_out) {
  GAI_CHECK_ARG(g != nullptr && rowptr_out_d != nullptr && colidx_out_d != nullptr && nnz_out != nullptr && (keep_ids_d != nullptr || n_keep == 0));
  cudaStream_t st = gai::S(stream);
  const uint32_t nv = gai_csr_nv(g);
  GAI_CHECK_ARG(n_keep <= nv);
  *rowptr_out_d = nullptr; *colidx_out_d = nullptr; *nnz_out = 0;
  uint32_t *new_id = nullptr, *deg = nullptr, *rp = nullptr, *ci = nullptr;
  void* tmp = nullptr;
  int rc = GAI_OK;
  do {
    if (cudaMalloc(&new_id, sizeof(uint32_t) * (nv ? nv : 1)) != cudaSuccess || cudaMalloc(&deg, sizeof(uint32_t) * ((size_t)n_keep + 1)) != cudaSuccess ||
        cudaMalloc(&rp, sizeof(uint32_t) * ((size_t)n_keep + 1)) != cudaSuccess) { cudaGetLastError(); rc = gai::set_error(GAI_ERR_NOMEM, "gai_induced_subgraph", "scratch"); break; }
    cudaMemsetAsync(new_id, 0xff, sizeof(uint32_t) * (nv ? nv : 1), st);
    cudaMemsetAsync(deg, 0, sizeof(uint32_t) * ((size_t)n_keep + 1), st);
    const unsigned warp_blocks = (unsigned)(((size_t)n_keep * 32 + 255) / 256);
    if (n_keep) {
      scatter_ids_kernel<<<(n_keep + 255) / 256, 256, 0, st>>>(n_keep, keep_ids_d, new_id);
      induced_rows_kernel<false><<<warp_blocks, 256, 0, st>>>(n_keep, keep_ids_d, gai_csr_rowptr(g), gai_csr_colidx(g), new_id, deg, nullptr, nullptr);
      __atomic_fetch_add(&gai::g_launches, 2ull, __ATOMIC_RELAXED);
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, deg, rp, (int)n_keep + 1, st);  // n_keep + 1 items: the last output is the total
    if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1) != cudaSuccess) { cudaGetLastError(); rc = gai::set_error(GAI_ERR_NOMEM, "gai_induced_subgraph", "scan workspace"); break; }
    if (cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, deg, rp, (int)n_keep + 1, st) != cudaSuccess) { rc = gai::set_error(GAI_ERR_CUDA, "gai_induced_subgraph", "scan"); break; }
    uint32_t total = 0;
    if (cudaMemcpyAsync(&total, rp + n_keep, sizeof(uint32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = gai::set_error(GAI_ERR_CUDA, "gai_induced_subgraph", cudaGetErrorString(cudaGetLastError())); break; }
    if (cudaMalloc(&ci, sizeof(uint32_t) * (total ? total : 1)) != cudaSuccess) { cudaGetLastError(); rc = gai::set_error(GAI_ERR_NOMEM, "gai_induced_subgraph", "columns"); break; }
    if (n_keep) {
      induced_rows_kernel<true><<<warp_blocks, 256, 0, st>>>(n_keep, keep_ids_d, gai_csr_rowptr(g), gai_csr_colidx(g), new_id, nullptr, rp, ci);
      __atomic_fetch_add(&gai::g_launches, 1ull, __ATOMIC_RELAXED);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { rc = gai::set_error(GAI_ERR_CUDA, "gai_induced_subgraph", cudaGetErrorString(cudaGetLastError())); break; }
    *rowptr_out_d = rp; *colidx_out_d = ci; *nnz_out = total;
    rp = nullptr; ci = nullptr;
  } while (0);
  cudaFree(new_id); cudaFree(deg); cudaFree(tmp); cudaFree(rp); cudaFree(ci);
  return rc;
}

int gai_add_selfloop_d(uint32_t nv, uint32_t first_id, const uint32_t* rowptr_d, const uint32_t* colidx_d, uint32_t* rowptr_out_d, uint32_t* colidx_out_d,
                       gai_stream_t stream) {
  GAI_CHECK_ARG(rowptr_d != nullptr && rowptr_out_d != nullptr && colidx_out_d != nullptr);
  GAI_CHECK_ARG(rowptr_d != rowptr_out_d && colidx_d != colidx_out_d);
  const size_t threads = ((size_t)nv + 1) * 32;
  selfloop_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, gai::S(stream)>>>(nv, first_id, rowptr_d, colidx_d, rowptr_out_d, colidx_out_d);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

}  // extern "C"
