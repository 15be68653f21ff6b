// Host-side mirror of the reference's graph + loader API for the GNN path, over the C ABI (include/gai_b200.h).
//
//   LearningGraph  <- include/gnn/lgraph.h:20-273   (host CSR; the device side is one gai_csr_t handle instead of the
//                                                     d_rowptr_/d_colidx_/d_vertex_data_/d_edge_data_ pointer set)
//   Reader         <- include/gnn/reader.h:4-40, src/gnn/reader.cpp:248-457 (graph.meta.txt / .vertex.bin / .edge.bin /
//                                                     .feats.bin / .vlabel.bin; DATASET_PATH; dataset whitelist)
// Same names, argument meaning and error behaviour (exit(1) on missing files / unknown dataset, exit(EXIT_FAILURE) on
// device errors) so that reference-side callers compile against it unchanged.
#pragma once
#include <cstdint>
#include <ostream>
#include <string>
#include <vector>
#include "gai_b200.h"

namespace gai_host {
class Comm;
void die_on(int status, const char* what);  // prints gai_last_error() and exits: the reference's CUDA_CHECK contract
gai_stream_t stream();                      // the stream every host-class call of THIS THREAD is issued on (one rank per thread)
void set_stream(gai_stream_t s);
// Log stream of the reference's progress lines (std::cout) — silenced per thread for the ranks > 0 of a partitioned run.
std::ostream& out();
void set_quiet(bool quiet);

// Per-op device timing (the reference's `time_ops` buckets, include/gnn/global.h:42-54, taken with CUDA events on the
// launching stream instead of gettimeofday around synchronous calls). Off by default; when on, every ABI call the
// host classes make is bracketed by an event pair tagged with its algorithmic bytes / flops.
void profile_enable(bool on);
bool profile_enabled();
struct OpScope {
  OpScope(const char* bucket, const std::string& shape, double bytes, double flops);
  ~OpScope();
  int idx;
};
std::string profile_collect_json();  // synchronises, aggregates per (bucket, shape), clears the records
}  // namespace gai_host

typedef uint32_t index_t;
typedef uint8_t label_t;
typedef uint8_t mask_t;
typedef uint8_t vlabel_t;
typedef float vdata_t;
typedef float edata_t;
typedef float acc_t;
typedef std::vector<float> vec_t;

class LearningGraph {
 public:
  typedef size_t iterator;
  explicit LearningGraph(bool use_gpu = true) : is_device(use_gpu) {}

  // construction (host)
  void allocateFrom(index_t nv, index_t ne);
  void fixEndEdge(index_t vid, index_t row_end) { rowptr_[vid + 1] = row_end; }
  void constructEdge(index_t eid, index_t dst) { colidx_[eid] = dst; }
  void add_selfloop();
  void degree_counting();
  LearningGraph* generate_masked_graph(mask_t* masks);

  // device residency
  void alloc_on_device();            // kept for API parity; allocation happens in copy_to_gpu
  void alloc_on_device(index_t n);
  void copy_to_gpu();                // uploads the CSR and builds norms / hub list / (lazily) the transpose permutation
  void compute_vertex_data();        // 1/sqrt(deg): done on the device inside copy_to_gpu; this re-derives it if the CSR changed
  void compute_edge_data();          // per-edge norms are never materialised (computed on the fly from vertex norms)
  void dealloc();

  // ---- 1D partition (one rank's rows; host/gai_dist.h) -----------------------------------------------------------------------
  // The graph holds the rows of the masters [first, first + size()) with GLOBAL column ids until partition_rows() renumbers them.
  void add_selfloop_rows(index_t first);   // add_selfloop on a block of rows: global id first + r enters row r at its sorted place
  // Column ids -> local ids: a master g becomes g - first, a remote neighbour becomes size() + its rank in the ascending list of
  // distinct remote neighbours (the halo). Edge order inside a row is untouched, so aggregated rows keep the single-GPU bits.
  void partition_rows(gai_host::Comm* comm, index_t nv_global);
  void partition_rows(int world, int rank, index_t nv_global);  // the integer part alone (no device, no peer group): tests
  bool partitioned() const { return comm_ != nullptr; }
  gai_host::Comm* comm() const { return comm_; }
  size_t rows_with_halo() const { return (size_t)num_vertices_ + halo_gids_.size(); }
  size_t num_halo() const { return halo_gids_.size(); }
  index_t first_global() const { return first_; }
  index_t size_global() const { return comm_ ? nv_global_ : num_vertices_; }
  const std::vector<index_t>& halo_global_ids() const { return halo_gids_; }
  // A matrix some aggregation gathers from holds this rank's master rows; it is registered once (collective, same order on every rank) ...
  void register_gather_buffer(const float* buf);
  // ... and before an aggregation reads it, the rows of the halo vertices are fetched from their owners' instances of the same matrix
  // (barrier - pull - barrier) into a scratch buffer shared by all exchanges of this graph; the returned pointer (row k = halo vertex k,
  // pitch ld) is what the aggregation reads neighbour ids >= size() from. NULL when there is nothing to exchange.
  const float* halo_exchange(const float* buf, int F, size_t ld);
  // The same into a caller-owned buffer of num_halo() x ld floats (halo rows that stay valid across steps: the input features).
  void halo_exchange_into(const float* buf, int F, size_t ld, float* dst);
  unsigned long long halo_exchanges = 0, halo_bytes = 0;  // counted per call (measurement)
  // Pipelined form (opt-in, GAI_HALO_BLOCKS > 1): aggregation is independent per feature column, so the exchange is cut into column
  // blocks (multiples of 32 columns) that cross NVLink on a second stream while the blocks already here are aggregated on the main one:
  // only the first block's transfer is exposed. Measured on the papers100M-shaped GCN (round 2, 2 and 8 GPUs): the exchange hides as
  // designed (HALO 14.5 -> 6.5 ms per epoch at N = 2) but the aggregation pays more than that for walking the CSR once per block
  // (75.7 -> 107.6 ms: at average degree 14 the per-row work, not the bytes, bounds it), so the one-piece exchange stays the default. begin() issues barrier (main stream) -> all block pulls + closing barrier (pull stream); wait_block(k) makes the main
  // stream wait for block k; end() makes it wait for the closing barrier (the owners may overwrite the matrix again). Every rank calls
  // the same sequence (a rank without halo takes part in the barriers). Results are bit-identical to the one-piece exchange.
  struct HaloBlocks {
    int n = 1;
    int col0[8] = {0}, ncol[8] = {0};
    const float* halo = nullptr;
  };
  static int halo_block_count(int F);   // GAI_HALO_BLOCKS (default 1 = one piece) capped so that a block keeps >= 64 columns
  HaloBlocks halo_exchange_begin(const float* buf, int F, size_t ld);
  void halo_wait_block(int k);
  void halo_exchange_end();

  size_t size() const { return num_vertices_; }
  size_t sizeEdges() const { return num_edges_; }
  bool on_device() const { return is_device; }
  index_t get_max_degree() const { return max_degree; }
  index_t get_degree(index_t v) const { return rowptr_[v + 1] - rowptr_[v]; }
  iterator begin() const { return 0; }
  iterator end() const { return num_vertices_; }

  // host accessors (lgraph.h:116-122)
  index_t* row_start_host_ptr() { return rowptr_.data(); }
  index_t* edge_dst_host_ptr() { return colidx_.data(); }
  index_t getEdgeDstHost(index_t eid) const { return colidx_[eid]; }
  index_t edge_begin_host(index_t vid) const { return rowptr_[vid]; }
  index_t edge_end_host(index_t vid) const { return rowptr_[vid + 1]; }

  // device accessors (lgraph.h:173-181)
  const index_t* row_start_ptr() const { return gai_csr_rowptr(dev_); }
  const index_t* edge_dst_ptr() const { return gai_csr_colidx(dev_); }
  const vdata_t* vertex_data_ptr() const { return gai_csr_vertex_norm(dev_); }
  gai_csr_t device() const { return dev_; }

 private:
  bool is_device;
  index_t num_vertices_ = 0, num_edges_ = 0, max_degree = 0;
  std::vector<index_t> rowptr_, colidx_;
  gai_csr_t dev_ = nullptr;
  gai_host::Comm* comm_ = nullptr;
  index_t nv_global_ = 0, first_ = 0;
  std::vector<index_t> halo_gids_;
  gai_halo_plan_t plan_ = nullptr;
  float* halo_scratch_ = nullptr;
  size_t halo_scratch_floats_ = 0;
  void ensure_halo_scratch(size_t floats);
  gai_stream_t pull_stream_ = nullptr;
  void* ev_fork_ = nullptr;
  void* ev_done_ = nullptr;
  void* ev_block_[8] = {nullptr};
};

typedef LearningGraph Graph;

class Reader {
 public:
  Reader() {}
  explicit Reader(std::string dataset) : dataset_str(dataset) {}
  void init(std::string dataset) { dataset_str = dataset; }

  // legacy text / .csgr dataset layout (reader.cpp:16-246): <DATASET_PATH><name>/<name>.csgr, <name>-labels.txt, <name>.ft or
  // <name>-feats.bin + <name>-dims.txt, <name>-{train,val,test}_mask.txt
  size_t csgr_read_labels(std::vector<label_t>& labels, bool is_single_class = true);
  size_t csgr_read_features(std::vector<float>& feats, std::string filetype = "bin");
  size_t csgr_read_masks(std::string mask_type, size_t n, size_t& begin, size_t& end, mask_t* masks);
  void csgr_read_graph(LearningGraph* g);

  void bin_read_graph(LearningGraph* g);
  size_t bin_read_features(std::vector<float>& feats);
  int bin_read_vlabels(std::vector<label_t>& labels, bool is_single_class = true);
  size_t bin_read_masks(std::string mask_type, size_t n, size_t& begin, size_t& end, mask_t* masks);

 private:
  std::string dataset_str, inputfile_path;
  index_t feat_len = 0, num_vertices_ = 0, num_edges_ = 0;
  int num_vertex_classes = 0, num_edge_classes = 0;
  int train_begin = 0, train_end = 0, train_count = 0;
  int val_begin = 0, val_end = 0, val_count = 0;
  int test_begin = 0, test_end = 0, test_count = 0;
};
