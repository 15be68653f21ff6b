/* gai_graph_gpu.cuh — the reference's device-graph accessor surface (class GraphGPU, include/graph_gpu.h:13-120,207-243) over the
 * device CSR of this library (gai_csr_t), for CUDA translation units that consume a graph the way the reference's non-GNN kernels do
 * (src/triangle/gpu_kernels/bs_warp_vertex.cuh, the traversal / PageRank kernels): pass the object BY VALUE into a kernel and call
 * V(), E(), N(v), N(v, e), get_degree(v), getOutDegree(v), edge_begin(v), edge_end(v), getEdgeDst(e), rowptr(), colidx().
 *
 * Differences from the reference class, all behind the same member names:
 *   - row offsets are 32-bit in device memory (the GNN path's CSR, nnz < 2^32); edge_begin / edge_end / E() still return eidType (int64);
 *     rowptr() therefore returns const uint32_t* (out_rowptr / in_rowptr alias it: the GNN graphs are symmetric);
 *   - the object is a VIEW: it owns nothing. init(gai_csr_t) borrows the arrays of a handle built by gai_csr_create /
 *     gai_csr_create_device (one rank's induced subgraph of a 1D partition included, gai_partition1d_h), init(nv, ne, rowptr_d, colidx_d)
 *     borrows caller arrays; release() forgets them. allocateFrom / copyToDevice / toHost are not reproduced — gai_csr_create uploads,
 *     gai_memcpy_d2h downloads.
 */
#ifndef GAI_GRAPH_GPU_CUH
#define GAI_GRAPH_GPU_CUH
#include <cstdint>
#include "gai_b200.h"

typedef uint32_t vidType;
typedef int64_t eidType;

class GraphGPU {
 protected:
  vidType num_vertices;
  eidType num_edges;
  int device_id, n_gpu;
  vidType max_degree;
  const uint32_t* d_rowptr;
  const vidType* d_colidx;

 public:
  GraphGPU(int n = 0, int m = 1) : num_vertices(0), num_edges(0), device_id(n), n_gpu(m), max_degree(0), d_rowptr(nullptr), d_colidx(nullptr) {}
  explicit GraphGPU(gai_csr_t g, int n = 0, int m = 1) : GraphGPU(n, m) { init(g); }
  void init(gai_csr_t g) {
    num_vertices = gai_csr_nv(g); num_edges = (eidType)gai_csr_nnz(g);
    d_rowptr = gai_csr_rowptr(g); d_colidx = gai_csr_colidx(g);
    max_degree = gai_csr_max_degree(g);
  }
  void init(gai_csr_t g, int n, int m) { device_id = n; n_gpu = m; init(g); }
  void init(vidType nv, eidType ne, const uint32_t* rowptr_d, const vidType* colidx_d, vidType max_deg = 0) {
    num_vertices = nv; num_edges = ne; d_rowptr = rowptr_d; d_colidx = colidx_d; max_degree = max_deg;
  }
  void release() { d_rowptr = nullptr; d_colidx = nullptr; num_vertices = 0; num_edges = 0; }
  inline __device__ __host__ bool is_directed() const { return false; }
  inline __device__ __host__ int get_num_devices() const { return n_gpu; }
  inline __device__ __host__ vidType V() const { return num_vertices; }
  inline __device__ __host__ vidType size() const { return num_vertices; }
  inline __device__ __host__ eidType E() const { return num_edges; }
  inline __device__ __host__ eidType sizeEdges() const { return num_edges; }
  inline __device__ __host__ vidType get_max_degree() const { return max_degree; }
  inline __device__ __host__ bool valid_vertex(vidType vertex) const { return vertex < num_vertices; }
  inline __device__ __host__ bool valid_edge(eidType edge) const { return edge < num_edges; }
  inline __device__ const vidType* N(vidType vid) const { return d_colidx + d_rowptr[vid]; }
  inline __device__ vidType N(vidType v, eidType e) const { return d_colidx[d_rowptr[v] + e]; }
  inline __device__ __host__ const uint32_t* rowptr() const { return d_rowptr; }
  inline __device__ __host__ const vidType* colidx() const { return d_colidx; }
  inline __device__ __host__ const uint32_t* out_rowptr() const { return d_rowptr; }
  inline __device__ __host__ const vidType* out_colidx() const { return d_colidx; }
  inline __device__ __host__ const uint32_t* in_rowptr() const { return d_rowptr; }
  inline __device__ __host__ const vidType* in_colidx() const { return d_colidx; }
  inline __device__ eidType getOutDegree(vidType src) const { return (eidType)(d_rowptr[src + 1] - d_rowptr[src]); }
  inline __device__ eidType getInDegree(vidType src) const { return getOutDegree(src); }
  inline __device__ vidType get_degree(vidType src) const { return d_rowptr[src + 1] - d_rowptr[src]; }
  inline __device__ vidType getEdgeDst(eidType edge) const { return d_colidx[edge]; }
  inline __device__ vidType getOutEdgeDst(eidType edge) const { return d_colidx[edge]; }
  inline __device__ vidType getInEdgeDst(eidType edge) const { return d_colidx[edge]; }
  inline __device__ eidType edge_begin(vidType src) const { return (eidType)d_rowptr[src]; }
  inline __device__ eidType edge_end(vidType src) const { return (eidType)d_rowptr[src + 1]; }
  inline __device__ eidType out_edge_begin(vidType src) const { return edge_begin(src); }
  inline __device__ eidType out_edge_end(vidType src) const { return edge_end(src); }
  inline __device__ eidType in_edge_begin(vidType src) const { return edge_begin(src); }
  inline __device__ eidType in_edge_end(vidType src) const { return edge_end(src); }
};

#endif /* GAI_GRAPH_GPU_CUH */
