// A consumer of the GraphGPU accessor surface (include/gai_graph_gpu.cuh) over this library's device CSR: the reference's vertex-parallel
// triangle kernel (src/triangle/gpu_kernels/bs_warp_vertex.cuh: one warp per vertex v, for every neighbour u the size of N(v) ∩ N(u) by
// binary search, include/operations.cuh intersect_num) restated on g.N(v) / g.getOutDegree(v). It is here to show — and test — that the
// device graph the GNN path builds (whole, or one rank's induced subgraph of the 1D partition, as src/triangle/multigpu_induced.cu uses it)
// serves the reference's non-GNN kernels through the accessors they are written against; it is not part of the GNN hot path.
#include "gai_graph_gpu.cuh"
#include <vector>
#include "gai_internal.cuh"

namespace {

// |a ∩ b| for two sorted lists, one warp: every lane searches its elements of the shorter list in the longer one
__device__ __forceinline__ unsigned long long intersect_num_warp(const vidType* a, vidType na, const vidType* b, vidType nb, int lane) {
  if (na > nb) { const vidType* t = a; a = b; b = t; const vidType tn = na; na = nb; nb = tn; }
  unsigned long long c = 0;
  for (vidType i = lane; i < na; i += 32) {
    const vidType key = a[i];
    vidType lo = 0, hi = nb;
    while (lo < hi) {
      const vidType mid = lo + ((hi - lo) >> 1);
      const vidType v = b[mid];
      if (v == key) { c++; break; }
      if (v < key) lo = mid + 1; else hi = mid;
    }
  }
  return c;
}

__global__ void triangle_warp_vertex_kernel(vidType begin, vidType end, GraphGPU g, unsigned long long* __restrict__ partial) {
  const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t num_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long count = 0;
  for (size_t v = warp_id + begin; v < end; v += num_warps) {
    const vidType* v_ptr = g.N((vidType)v);
    const vidType v_size = (vidType)g.getOutDegree((vidType)v);
    for (vidType e = 0; e < v_size; e++) {
      const vidType u = v_ptr[e];
      count += intersect_num_warp(v_ptr, v_size, g.N(u), (vidType)g.getOutDegree(u), lane);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
  if (lane == 0) partial[warp_id] = count;  // one slot per warp, summed on the host: integer, order-free, no atomics
}

}  // namespace

extern "C" {

// sum over v in [begin, end) and u in N(v) of |N(v) ∩ N(u)| (= 6 x the triangles of a symmetric graph when the range is everything)
int gai_triangle_count_rows(gai_csr_t g, uint32_t begin, uint32_t end, uint64_t* count_h, gai_stream_t stream) {
  GAI_CHECK_ARG(g != nullptr && count_h != nullptr && begin <= end && end <= gai_csr_nv(g));
  *count_h = 0;
  if (begin == end) return GAI_OK;
  cudaStream_t st = gai::S(stream);
  const unsigned blocks = (unsigned)gai::sm_count() * 8, threads = 256;
  const size_t warps = (size_t)blocks * threads / 32;
  void* ws = nullptr;
  int rc = gai::workspace(sizeof(unsigned long long) * warps, &ws, st);
  if (rc != GAI_OK) return rc;
  GraphGPU view(g);
  triangle_warp_vertex_kernel<<<blocks, threads, 0, st>>>(begin, end, view, reinterpret_cast<unsigned long long*>(ws));
  GAI_LAUNCH_CHECK();
  std::vector<unsigned long long> h(warps);
  GAI_CUDA(cudaMemcpyAsync(h.data(), ws, sizeof(unsigned long long) * warps, cudaMemcpyDeviceToHost, st));
  GAI_CUDA(cudaStreamSynchronize(st));
  unsigned long long total = 0;
  for (unsigned long long x : h) total += x;
  *count_h = total;
  return GAI_OK;
}

}  // extern "C"
