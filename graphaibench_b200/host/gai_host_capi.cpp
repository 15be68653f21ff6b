// C handles over the host classes, for ctypes callers (tests, bench.py). Not part of the reference boundary: the
// reference-facing surface is the C++ API in gai_graph.h / gai_layers.h / gai_model.h.
#include <cstring>
#include <thread>
#include "gai_converter.h"
#include "gai_model.h"
#include "gai_sampler.h"

namespace {
struct ModelBase {
  virtual ~ModelBase() {}
  virtual float train_epoch(float* loss) = 0;
  virtual float forward(float* loss) = 0;
  virtual void backward() = 0;
  virtual void update() = 0;
  virtual float evaluate(const char* which) = 0;
  virtual void refresh(const float* feats) = 0;
  virtual void prefetch(const float* feats) = 0;
  // device pointer + logical element count; cols/ld != 0 for per-vertex tensors stored with a row pitch
  virtual float* tensor(const char* name, int layer, size_t* n, size_t* cols, size_t* ld) = 0;
  virtual Graph* graph() = 0;
};
template <typename L>
struct Box : ModelBase {
  Model<L> m;
  float train_epoch(float* loss) override { return m.train_epoch(*loss); }
  float forward(float* loss) override { m.set_netphases(net_phase::TRAIN); return m.forward_prop(*loss); }
  void backward() override { m.backward_prop(); }
  void update() override { m.update_weights(m.shared_optimizer()); }
  float evaluate(const char* which) override { return m.evaluate(which); }
  void refresh(const float* feats) override { m.refresh_inputs_from_host(feats); }
  void prefetch(const float* feats) override { m.prefetch_features_from_host(feats); }
  Graph* graph() override { return m.graph(); }
  float* extra(GCN_layer&, const std::string&, size_t*) { return nullptr; }
  float* extra(SAGE_layer&, const std::string&, size_t*) { return nullptr; }
  float* extra(GAT_layer& y, const std::string& name, size_t* n) {
    GAT_Aggregator& a = y.aggregator_ref();
    *n = y.get_dim_out();
    if (name == "alpha_l") return a.d_alpha_l;
    if (name == "alpha_r") return a.d_alpha_r;
    if (name == "alpha_lgrad") return a.d_alpha_lgrad;
    if (name == "alpha_rgrad") return a.d_alpha_rgrad;
    return nullptr;
  }
  float* tensor(const char* name_, int l, size_t* n, size_t* cols, size_t* ld) override {
    std::string name(name_);
    *cols = 0; *ld = 0;
    if (name == "dense_W" && m.dense()) { *n = (size_t)m.dense()->dim_in * m.dense()->dim_out; return m.dense()->d_weight; }
    if (name == "dense_W_grad" && m.dense()) { *n = (size_t)m.dense()->dim_in * m.dense()->dim_out; return m.dense()->d_weight_grad; }
    if (l < 0 || l >= m.num_conv_layers()) return nullptr;
    L& y = m.conv_layer(l);
    float* p = y.weight_ptr(name);
    if (p) { *n = y.weight_size(name); y.tensor_layout(name, cols, ld); return p; }
    return extra(y, name, n);
  }
};
}  // namespace

extern "C" {

void gai_host_set_stream(void* s) { gai_host::set_stream(s); }

// Glorot initial weights exactly as the layers draw them (init_glorot, gai_layers.cpp): used by the 1D-partitioned trainer,
// whose replicated weights must start from the reference's values.
void gai_host_glorot(uint64_t dim_x, uint64_t dim_y, unsigned seed, float* out_h) {
  vec_t w;
  init_glorot(dim_x, dim_y, w, seed);
  memcpy(out_h, w.data(), sizeof(float) * w.size());
}

void* gai_graph_new(uint32_t nv, uint32_t ne, const uint32_t* rowptr, const uint32_t* colidx) {
  Graph* g = new Graph(true);
  g->allocateFrom(nv, ne);
  for (uint32_t v = 0; v < nv; v++) g->fixEndEdge(v, rowptr[v + 1]);
  memcpy(g->edge_dst_host_ptr(), colidx, sizeof(uint32_t) * ne);
  return g;
}

// arch: 0 GCN, 1 SAGE, 2 GAT. Takes ownership of the graph (self-loops are added inside, as Model::load_data does).
void* gai_model_new(int arch, void* graph, int dim_init, int dim_hid, int num_cls, int num_layers, float lr, const float* feats_h,
                    const uint8_t* labels_h, const int64_t* split9) {
  Graph* g = (Graph*)graph;
  if (arch == 0) { auto* b = new Box<GCN_layer>(); b->m.init_from_memory(gnn_arch::GCN, g, dim_init, num_cls, feats_h, labels_h, split9, dim_hid, num_layers, lr, 0, 1 << 30); b->m.construct_network(); return (ModelBase*)b; }
  if (arch == 1) { auto* b = new Box<SAGE_layer>(); b->m.init_from_memory(gnn_arch::SAGE, g, dim_init, num_cls, feats_h, labels_h, split9, dim_hid, num_layers, lr, 0, 1 << 30); b->m.construct_network(); return (ModelBase*)b; }
  if (arch == 2) { auto* b = new Box<GAT_layer>(); b->m.init_from_memory(gnn_arch::GAT, g, dim_init, num_cls, feats_h, labels_h, split9, dim_hid, num_layers, lr, 0, 1 << 30); b->m.construct_network(); return (ModelBase*)b; }
  return nullptr;
}
float gai_model_train_epoch(void* m, float* loss) { return ((ModelBase*)m)->train_epoch(loss); }
float gai_model_forward(void* m, float* loss) { return ((ModelBase*)m)->forward(loss); }
void gai_model_backward(void* m) { ((ModelBase*)m)->backward(); }
void gai_model_update(void* m) { ((ModelBase*)m)->update(); }
float gai_model_evaluate(void* m, const char* which) { return ((ModelBase*)m)->evaluate(which); }
void gai_model_refresh_inputs(void* m, const float* feats_h) { ((ModelBase*)m)->refresh(feats_h); }
void gai_model_prefetch_features(void* m, const float* feats_h) { ((ModelBase*)m)->prefetch(feats_h); }
int64_t gai_model_tensor_size(void* m, const char* name, int layer) {
  size_t n = 0, cols = 0, ld = 0;
  return ((ModelBase*)m)->tensor(name, layer, &n, &cols, &ld) ? (int64_t)n : -1;
}
// Dense host copies of a named tensor; pitched per-vertex tensors are packed / unpacked row by row.
int64_t gai_model_get(void* m, const char* name, int layer, float* out_h, int64_t cap) {
  size_t n = 0, cols = 0, ld = 0;
  float* p = ((ModelBase*)m)->tensor(name, layer, &n, &cols, &ld);
  if (!p || (int64_t)n > cap) return -1;
  if (cols && ld != cols) {
    gai_host::die_on(gai_memcpy2d(out_h, cols * sizeof(float), p, ld * sizeof(float), cols * sizeof(float), n / cols, gai_host::stream()), "gai_memcpy2d");
    gai_host::die_on(gai_stream_sync(gai_host::stream()), "gai_stream_sync");
  } else {
    copy_float_to_host(n, p, out_h);
  }
  return (int64_t)n;
}
int64_t gai_model_set(void* m, const char* name, int layer, const float* in_h, int64_t n_in) {
  size_t n = 0, cols = 0, ld = 0;
  float* p = ((ModelBase*)m)->tensor(name, layer, &n, &cols, &ld);
  if (!p || (int64_t)n != n_in) return -1;
  if (cols && ld != cols) {
    gai_host::die_on(gai_memcpy2d(p, ld * sizeof(float), in_h, cols * sizeof(float), cols * sizeof(float), n / cols, gai_host::stream()), "gai_memcpy2d");
    gai_host::die_on(gai_stream_sync(gai_host::stream()), "gai_stream_sync");
  } else {
    copy_float_to_device(n, in_h, p);
  }
  return (int64_t)n;
}
void gai_host_profile_enable(int on) { gai_host::profile_enable(on != 0); }
// Writes the aggregated per-op timings as JSON into buf (NUL-terminated, truncated to cap); returns the full length.
int64_t gai_host_profile_json(char* buf, int64_t cap) {
  const std::string js = gai_host::profile_collect_json();
  if (buf && cap > 0) { const size_t n = js.size() < (size_t)cap - 1 ? js.size() : (size_t)cap - 1; memcpy(buf, js.data(), n); buf[n] = 0; }
  return (int64_t)js.size();
}
// Reader (src/gnn/reader.cpp:248-457 mirror) on $DATASET_PATH/<dataset>/: same two-call protocol and meta layout as the reference-side
// harness (oracle/ref_harness.cpp: ref_reader_load), so a test can compare the two loaders field by field. Host only, no device work.
int gai_reader_load(const char* dataset, int single_class, int64_t* meta, uint32_t* rowptr, uint32_t* colidx, float* feats, uint8_t* labels) {
  Reader reader{std::string(dataset)};
  Graph g(true);
  reader.bin_read_graph(&g);
  std::vector<float> f;
  const size_t flen = reader.bin_read_features(f);
  std::vector<label_t> lab;
  const int ncls = reader.bin_read_vlabels(lab, single_class != 0);
  size_t b[3], e[3], c[3];
  const char* kinds[3] = {"train", "val", "test"};
  for (int i = 0; i < 3; i++) c[i] = reader.bin_read_masks(kinds[i], g.size(), b[i], e[i], nullptr);
  meta[0] = (int64_t)g.size(); meta[1] = (int64_t)g.sizeEdges(); meta[2] = (int64_t)flen; meta[3] = ncls;
  for (int i = 0; i < 3; i++) { meta[4 + 3 * i] = (int64_t)b[i]; meta[5 + 3 * i] = (int64_t)e[i]; meta[6 + 3 * i] = (int64_t)c[i]; }
  if (rowptr) {
    memcpy(rowptr, g.row_start_host_ptr(), sizeof(uint32_t) * (g.size() + 1));
    memcpy(colidx, g.edge_dst_host_ptr(), sizeof(uint32_t) * g.sizeEdges());
    memcpy(feats, f.data(), sizeof(float) * f.size());
    memcpy(labels, lab.data(), lab.size());
  }
  return 0;
}
// ---- 1D-partitioned training (host/gai_dist.h) --------------------------------------------------------------------------------
void* gai_comm_new(int rank, int world, gai_allgather_fn allgather, void* ctx) { return new gai_host::Comm(rank, world, allgather, ctx); }
void gai_comm_free(void* c) { delete (gai_host::Comm*)c; }
void gai_comm_barrier(void* c) { ((gai_host::Comm*)c)->barrier(); }
void gai_comm_check(void* c) { ((gai_host::Comm*)c)->check(); }

// One rank's model: its rows of the RAW graph (global column ids, rows_rowptr rebased to 0), its rows of features / labels, the GLOBAL split.
void* gai_model_new_partitioned(int arch, void* comm, uint32_t nv_global, const int64_t* rows_rowptr, const uint32_t* rows_colidx, int dim_init,
                                int dim_hid, int num_cls, int num_layers, float lr, const float* feats_local, const uint8_t* labels_local,
                                const int64_t* split9_global) {
  gai_host::Comm* c = (gai_host::Comm*)comm;
  if (arch == 0) { auto* b = new Box<GCN_layer>(); b->m.set_comm(c); b->m.init_partitioned(gnn_arch::GCN, nv_global, rows_rowptr, rows_colidx, dim_init, num_cls, feats_local, labels_local, split9_global, dim_hid, num_layers, lr); b->m.construct_network(); return (ModelBase*)b; }
  if (arch == 1) { auto* b = new Box<SAGE_layer>(); b->m.set_comm(c); b->m.init_partitioned(gnn_arch::SAGE, nv_global, rows_rowptr, rows_colidx, dim_init, num_cls, feats_local, labels_local, split9_global, dim_hid, num_layers, lr); b->m.construct_network(); return (ModelBase*)b; }
  return nullptr;  // GAT: single-GPU only
}
// out4 = {masters, halo rows, halo exchanges so far, halo bytes received so far}
void gai_model_halo_stats(void* m, uint64_t* out4) {
  Graph* g = ((ModelBase*)m)->graph();
  out4[0] = g->size(); out4[1] = g->num_halo(); out4[2] = g->halo_exchanges; out4[3] = g->halo_bytes;
}

// The whole partitioned path inside one process: `world` host threads = `world` ranks on the visible devices (rank r on device
// r mod device count), each taking its slice of the full host graph, training `epochs` epochs. Rank 0's per-epoch loss / accuracy and
// final weights come back ("W" / "W_self" of layer l at w_out + offsets the caller computes: layers in order, W then W_self).
int gai_host_train_partitioned(int arch, int world, uint32_t nv, const int64_t* rowptr64, const uint32_t* colidx, int dim_init, int dim_hid,
                               int num_cls, int num_layers, float lr, const float* feats, const uint8_t* labels, const int64_t* split9, int epochs,
                               float* losses_out, float* accs_out, float* test_acc_out, float* w_out, uint64_t* halo_stats_out /* world x 4 */) {
  int ndev = 0;
  if (gai_device_count(&ndev) != GAI_OK || ndev < 1) return -1;
  gai_host::ThreadGroup group(world);
  std::vector<gai_host::ThreadRank> ctx((size_t)world);
  std::vector<std::thread> threads;
  for (int r = 0; r < world; r++) {
    ctx[r] = {&group, r};
    threads.emplace_back([&, r]() {
      gai_host::die_on(gai_set_device(r % ndev), "gai_set_device");
      gai_stream_t st = nullptr;
      gai_host::die_on(gai_stream_create(&st), "gai_stream_create");
      gai_host::set_stream(st);
      gai_host::Comm* comm = new gai_host::Comm(r, world, gai_host::thread_allgather, &ctx[r]);
      const gai_host::OwnerRange own = gai_host::owner_range(nv, world, r);
      std::vector<int64_t> rp(own.last - own.first + 1);
      for (uint32_t v = own.first; v <= own.last; v++) rp[v - own.first] = rowptr64[v] - rowptr64[own.first];
      ModelBase* m = (ModelBase*)gai_model_new_partitioned(arch, comm, nv, rp.data(), colidx + rowptr64[own.first], dim_init, dim_hid, num_cls, num_layers, lr,
                                                           feats + (size_t)own.first * dim_init, labels + own.first, split9);
      for (int ep = 0; ep < epochs; ep++) {
        float loss = 0.f;
        const float acc = m->train_epoch(&loss);
        if (r == 0) { losses_out[ep] = loss; accs_out[ep] = acc; }
      }
      const float tacc = m->evaluate("test");
      comm->check();
      if (r == 0) {
        *test_acc_out = tacc;
        size_t off = 0;
        for (int l = 0; l < num_layers && w_out; l++) {
          for (const char* name : {"W", "W_self"}) {
            size_t n = 0, cols = 0, ld = 0;
            float* p = m->tensor(name, l, &n, &cols, &ld);
            if (p && n) { copy_float_to_host(n, p, w_out + off); off += n; }
          }
        }
      }
      if (halo_stats_out) gai_model_halo_stats(m, halo_stats_out + 4 * r);
      gai_stream_sync(st);
    });
  }
  for (auto& t : threads) t.join();
  return 0;
}
// Host-only (integer) part of the partition, for tests: this rank's rows (global column ids; self-loops added first if `selfloops`) ->
// local column ids + the halo list. Two-call protocol: halo_out == NULL returns the sizes only. sizes = {n_halo, nnz_out}.
int gai_host_partition_rows(int world, int rank, uint32_t nv_global, const int64_t* rows_rowptr, const uint32_t* rows_colidx, int selfloops,
                            uint64_t* sizes, uint32_t* rowptr_out, uint32_t* colidx_out, uint32_t* halo_out) {
  const gai_host::OwnerRange own = gai_host::owner_range(nv_global, world, rank);
  const uint32_t n = own.last - own.first;
  Graph g(false);
  g.allocateFrom(n, (index_t)rows_rowptr[n]);
  for (uint32_t v = 0; v < n; v++) g.fixEndEdge(v, (index_t)rows_rowptr[v + 1]);
  std::copy(rows_colidx, rows_colidx + rows_rowptr[n], g.edge_dst_host_ptr());
  if (selfloops) g.add_selfloop_rows(own.first);
  g.partition_rows(world, rank, nv_global);
  sizes[0] = g.num_halo(); sizes[1] = g.sizeEdges();
  if (halo_out) {
    std::copy(g.row_start_host_ptr(), g.row_start_host_ptr() + n + 1, rowptr_out);
    std::copy(g.edge_dst_host_ptr(), g.edge_dst_host_ptr() + g.sizeEdges(), colidx_out);
    std::copy(g.halo_global_ids().begin(), g.halo_global_ids().end(), halo_out);
  }
  return 0;
}
// Converter (host/gai_converter.h) for ctypes callers: file -> files, and pairs -> CSR (two-call: rowptr_out == NULL returns nv/ne in sizes).
int gai_convert_file(const char* file_type, const char* infile, const char* out_prefix, int is_bipartite) {
  Converter c(file_type, infile, is_bipartite != 0);
  c.generate_binary_graph(out_prefix, true, true, false, false);
  c.write_meta(out_prefix);
  return 0;
}
int gai_convert_pairs(int64_t nv, const uint32_t* src, const uint32_t* dst, uint64_t n, int symmetrize, int64_t* sizes, int64_t* rowptr_out, uint32_t* colidx_out) {
  static thread_local Converter* last = nullptr;
  if (!rowptr_out) {
    delete last;
    last = new Converter();
    last->from_pairs(nv, src, dst, (size_t)n, symmetrize != 0);
    sizes[0] = last->V(); sizes[1] = last->E();
    return 0;
  }
  if (!last) return -1;
  std::copy(last->row_offsets().begin(), last->row_offsets().end(), rowptr_out);
  std::copy(last->column_indices().begin(), last->column_indices().end(), colidx_out);
  delete last; last = nullptr;
  return 0;
}
// The legacy loader (csgr_* methods) with the same two-call protocol; meta = {nv, ne, feat_len, num_classes, then (begin, end, count) of
// train / val / test}; masks_out = 3 x nv bytes.
int gai_reader_load_csgr(const char* dataset, int single_class, int64_t* meta, uint32_t* rowptr, uint32_t* colidx, float* feats, uint8_t* labels,
                         uint8_t* masks_out) {
  Reader reader{std::string(dataset)};
  Graph g(true);
  reader.csgr_read_graph(&g);
  std::vector<float> f;
  const size_t flen = reader.csgr_read_features(f, "bin");
  std::vector<label_t> lab;
  const size_t ncls = reader.csgr_read_labels(lab, single_class != 0);
  std::vector<mask_t> masks(3 * g.size(), 0);
  size_t b[3], e[3], c[3];
  const char* kinds[3] = {"train", "val", "test"};
  for (int i = 0; i < 3; i++) c[i] = reader.csgr_read_masks(kinds[i], g.size(), b[i], e[i], masks.data() + (size_t)i * g.size());
  meta[0] = (int64_t)g.size(); meta[1] = (int64_t)g.sizeEdges(); meta[2] = (int64_t)flen; meta[3] = (int64_t)ncls;
  for (int i = 0; i < 3; i++) { meta[4 + 3 * i] = (int64_t)b[i]; meta[5 + 3 * i] = (int64_t)e[i]; meta[6 + 3 * i] = (int64_t)c[i]; }
  if (rowptr) {
    memcpy(rowptr, g.row_start_host_ptr(), sizeof(uint32_t) * (g.size() + 1));
    memcpy(colidx, g.edge_dst_host_ptr(), sizeof(uint32_t) * g.sizeEdges());
    memcpy(feats, f.data(), sizeof(float) * f.size());
    memcpy(labels, lab.data(), lab.size());
    memcpy(masks_out, masks.data(), masks.size());
  }
  return 0;
}
// Sampler (host/gai_sampler.h) on an in-memory graph, same protocol as the reference-side harness (oracle/ref_harness.cpp:
// ref_sampler_run). with_subgraph = 0 stops after select_vertices (host only, no device needed).
int64_t gai_sampler_run(uint32_t nv, const uint32_t* rowptr, const uint32_t* colidx, const uint8_t* masks_train, size_t count, uint32_t n, unsigned seed,
                        int with_subgraph, int64_t* sizes, uint32_t* set_out, uint32_t* rowptr_out, uint32_t* colidx_out) {
  Graph full(true);
  full.allocateFrom(nv, rowptr[nv]);
  for (uint32_t v = 0; v < nv; v++) full.fixEndEdge(v, rowptr[v + 1]);
  std::copy(colidx, colidx + rowptr[nv], full.edge_dst_host_ptr());
  gai_host::set_quiet(true);
  Graph* tg = full.generate_masked_graph(const_cast<mask_t*>(masks_train));
  Sampler sampler(&full, tg, const_cast<mask_t*>(masks_train), count);
  VertexSet st;
  sampler.select_vertices(n, st, seed);
  sizes[0] = (int64_t)st.size(); sizes[1] = 0;
  Graph sg(true);
  if (with_subgraph) {
    full.copy_to_gpu();
    std::vector<mask_t> masks(nv, 0);
    sampler.generateSubgraph(st, masks.data(), &sg);
    sizes[1] = (int64_t)sg.sizeEdges();
  }
  if (set_out) {
    size_t i = 0;
    for (auto v : st) set_out[i++] = v;
    if (with_subgraph) {
      std::copy(sg.row_start_host_ptr(), sg.row_start_host_ptr() + sg.size() + 1, rowptr_out);
      std::copy(sg.edge_dst_host_ptr(), sg.edge_dst_host_ptr() + sg.sizeEdges(), colidx_out);
    }
  }
  full.dealloc();
  delete tg;
  gai_host::set_quiet(false);
  return (int64_t)st.size();
}
void gai_model_sync() { gai_stream_sync(gai_host::stream()); }
}
