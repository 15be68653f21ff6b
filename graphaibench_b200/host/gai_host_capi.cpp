// C handles over the host classes, for ctypes callers (tests, bench.py). Not part of the reference boundary: the
// reference-facing surface is the C++ API in gai_graph.h / gai_layers.h / gai_model.h.
#include <cstring>
#include "gai_model.h"

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
void gai_model_sync() { gai_stream_sync(gai_host::stream()); }
}
