// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// C-callable harness around the UNMODIFIED reference CPU implementation (chenxuhao/GraphAIBench,
// src/gnn + src/layers + src/utilities). build_ref.sh compiles the reference sources where they lie
// (a patched scratch copy, see that script) together with this file into oracle/_ref/libref_gnn.so.
// Everything numeric below is executed by reference code; this file only wires buffers:
//   * graphs are built with LearningGraph::allocateFrom/constructEdge/fixEndEdge (include/gnn/lgraph.h:96-114)
//   * models are the reference's Model<L> (include/gnn/net.h:9-83) with load_data() bypassed (it reads files
//     and parses argv, src/gnn/net.cpp:12-204); the fields load_data would set are filled from the arguments,
//     then construct_network/forward_prop/backward_prop/update_weights/evaluate run as in Model::train
//     (src/gnn/net.cpp:361-419).
// Compiled with -fno-access-control so private members can be read for golden dumps.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "net.h"
#include "math_functions.hh"
#include "softmax_loss_layer.h"
#include "sigmoid_loss_layer.h"
#include "reader.h"
#include "sampler.h"

std::map<char, double> time_ops;  // reference global (defined in train.cpp:3, which is not linked here)

namespace {

template <typename L> struct arch_of;
template <> struct arch_of<GCN_layer> { static constexpr gnn_arch value = gnn_arch::GCN; };
template <> struct arch_of<SAGE_layer> { static constexpr gnn_arch value = gnn_arch::SAGE; };
template <> struct arch_of<GAT_layer> { static constexpr gnn_arch value = gnn_arch::GAT; };

struct ModelBase {
  virtual ~ModelBase() {}
  virtual float train_epoch(float* loss) = 0;
  virtual float forward_only(float* loss) = 0;
  virtual void backward_only() = 0;
  virtual void update_only() = 0;
  virtual float evaluate(const char* type) = 0;
  virtual int64_t tensor(const char* name, int layer, float** ptr) = 0;
};

template <typename L>
struct ModelBox : ModelBase {
  Model<L>* m;
  optimizer* opt;
  ModelBox(Graph* g, int nv, int dim_init, int dim_hid, int num_cls, int num_layers, float lr,
           const float* feats, const uint8_t* labels, const int64_t* split, int threads, bool sigmoid = false) {
    omp_set_num_threads(threads);
    openblas_set_num_threads(threads);
    m = (Model<L>*)operator new(sizeof(Model<L>));
    memset((void*)m, 0, sizeof(Model<L>));
    // placement-construct the non-trivial members that the zero fill does not make valid
    new (&m->dataset_name) std::string("harness");
    new (&m->layer_gconv) std::vector<L>();
    new (&m->input_features) std::vector<float>(feats, feats + (size_t)nv * dim_init);
    if (sigmoid) {  // multi-hot [nv x num_cls] from the class ids, as Reader::bin_read_vlabels(labels, false) (reader.cpp:347-412)
      new (&m->labels) std::vector<label_t>((size_t)nv * num_cls, 0);
      for (int v = 0; v < nv; v++) if (labels[v] < num_cls) m->labels[(size_t)v * num_cls + labels[v]] = 1;
    } else {
      new (&m->labels) std::vector<label_t>(labels, labels + nv);
    }
    new (&m->masks_train) std::vector<mask_t>(nv, 0);
    new (&m->masks_test) std::vector<mask_t>(nv, 0);
    new (&m->masks_val) std::vector<mask_t>(nv, 0);
    new (&m->feats_subg) vec_t();
    new (&m->labels_subg) std::vector<label_t>();
    new (&m->subgs) std::vector<Graph*>();
    m->num_epochs = 0; m->num_layers = num_layers; m->num_samples = nv; m->num_threads = threads;
    m->num_cls = num_cls; m->dim_init = dim_init; m->dim_hid = dim_hid; m->subg_size = 0; m->subg_nv = 0;
    m->val_interval = 1 << 30; m->feat_drop = 0.f; m->score_drop = 0.f; m->lrate = lr;
    m->train_begin = split[0]; m->train_end = split[1]; m->train_count = split[2];
    m->val_begin = split[3]; m->val_end = split[4]; m->val_count = split[5];
    m->test_begin = split[6]; m->test_end = split[7]; m->test_count = split[8];
    m->is_sigmoid = sigmoid; m->use_gpu = false; m->inductive = false;
    m->arch = arch_of<L>::value;
    // src/gnn/net.cpp:67-71
    m->use_l2norm = (m->arch == gnn_arch::GAT);
    m->use_dense = m->use_l2norm;
    // src/gnn/net.cpp:126-142
    for (int64_t i = split[0]; i < split[1]; i++) m->masks_train[i] = 1;
    for (int64_t i = split[3]; i < split[4]; i++) m->masks_val[i] = 1;
    for (int64_t i = split[6]; i < split[7]; i++) m->masks_test[i] = 1;
    m->full_graph = g;
    if (m->arch != gnn_arch::SAGE) g->add_selfloop();  // src/gnn/net.cpp:96
    m->training_graph = g;
    g->compute_vertex_data();  // src/gnn/net.cpp:199-202 (CPU, no MKL)
    m->construct_network();
    opt = new adam(lr);  // src/gnn/net.cpp:362
  }
  float forward_only(float* loss) override {
    m->set_netphases(net_phase::TRAIN);
    acc_t l = 0;
    float acc = m->forward_prop(l);
    *loss = l;
    return acc;
  }
  void backward_only() override { m->backward_prop(); }
  void update_only() override { m->update_weights(opt); }
  float train_epoch(float* loss) override {  // src/gnn/net.cpp:373-383
    float acc = forward_only(loss);
    m->backward_prop();
    m->update_weights(opt);
    return acc;
  }
  float evaluate(const char* type) override { return m->evaluate(std::string(type)); }
  int64_t tensor(const char* name_, int l, float** ptr) override {
    std::string name(name_);
    int64_t nv = m->num_samples;
    if (name == "logits") { *ptr = m->layer_loss->get_feat_in(); return nv * m->num_cls; }
    if (name == "probs") { *ptr = m->layer_loss->get_feat_out(); return nv * m->num_cls; }
    if (name == "l2norm_feat_in" && m->use_l2norm) { *ptr = m->layer_l2norm->feat_in; return nv * m->dim_hid; }
    if (name == "l2norm_grad_in" && m->use_l2norm) { *ptr = m->layer_l2norm->grad_in; return nv * m->dim_hid; }
    if (name == "dense_feat_in" && m->use_dense) { *ptr = m->layer_dense->feat_in; return nv * m->dim_hid; }
    if (name == "dense_grad_in" && m->use_dense) { *ptr = m->layer_dense->grad_in; return nv * m->num_cls; }
    if (name == "dense_W" && m->use_dense) { *ptr = m->layer_dense->weight.data(); return m->layer_dense->weight.size(); }
    if (name == "dense_W_grad" && m->use_dense) { *ptr = m->layer_dense->weight_grad.data(); return m->layer_dense->weight_grad.size(); }
    if (l < 0 || l >= m->num_layers) return -1;
    L& y = m->layer_gconv[l];
    if (name == "feat_in") { *ptr = y.feat_in; return nv * y.dim_in; }
    if (name == "grad_in") { *ptr = y.grad_in; return nv * y.dim_out; }
    if (name == "W") { *ptr = y.W_neigh.data(); return y.W_neigh.size(); }
    if (name == "W_grad") { *ptr = y.W_neigh_grad.data(); return y.W_neigh_grad.size(); }
    if (name == "W_self") { *ptr = y.W_self.data(); return y.W_self.size(); }
    if (name == "W_self_grad") { *ptr = y.W_self_grad.data(); return y.W_self_grad.size(); }
    if (name == "out_temp") { *ptr = y.out_temp.data(); return y.out_temp.size(); }
    if (name == "in_temp1") { *ptr = y.in_temp1.data(); return y.in_temp1.size(); }
    return extra(y, name, ptr);
  }
  int64_t extra(GCN_layer&, const std::string&, float**) { return -1; }
  int64_t extra(SAGE_layer&, const std::string&, float**) { return -1; }
  int64_t extra(GAT_layer& y, const std::string& name, float** ptr) {
    if (name == "alpha_l") { *ptr = y.aggr.alpha_l.data(); return y.aggr.alpha_l.size(); }
    if (name == "alpha_r") { *ptr = y.aggr.alpha_r.data(); return y.aggr.alpha_r.size(); }
    if (name == "alpha_lgrad") { *ptr = y.aggr.alpha_lgrad.data(); return y.aggr.alpha_lgrad.size(); }
    if (name == "alpha_rgrad") { *ptr = y.aggr.alpha_rgrad.data(); return y.aggr.alpha_rgrad.size(); }
    if (name == "norm_scores") { *ptr = y.aggr.norm_scores.data(); return y.aggr.norm_scores.size(); }
    if (name == "temp_scores") { *ptr = y.aggr.temp_scores.data(); return y.aggr.temp_scores.size(); }
    if (name == "scores") { *ptr = y.aggr.scores.data(); return y.aggr.scores.size(); }
    if (name == "norm_scores_grad") { *ptr = y.aggr.norm_scores_grad.data(); return y.aggr.norm_scores_grad.size(); }
    return -1;
  }
};

}  // namespace

extern "C" {

void* ref_graph_new(uint32_t nv, uint32_t ne, const uint32_t* rowptr, const uint32_t* colidx) {
  Graph* g = new Graph(false);
  g->allocateFrom(nv, ne);
  for (uint32_t v = 0; v < nv; v++) g->fixEndEdge(v, rowptr[v + 1]);
  for (uint32_t e = 0; e < ne; e++) g->constructEdge(e, colidx[e]);
  return g;
}
void ref_graph_add_selfloop(void* g) { ((Graph*)g)->add_selfloop(); }
void ref_graph_compute_vertex_data(void* g) { ((Graph*)g)->compute_vertex_data(); }
void ref_graph_compute_edge_data(void* g) { ((Graph*)g)->compute_edge_data(); }
uint32_t ref_graph_nv(void* g) { return ((Graph*)g)->size(); }
uint32_t ref_graph_ne(void* g) { return ((Graph*)g)->sizeEdges(); }
void ref_graph_export(void* g_, uint32_t* rowptr, uint32_t* colidx, float* vdata, float* edata) {
  Graph* g = (Graph*)g_;
  memcpy(rowptr, g->row_start_host_ptr(), sizeof(uint32_t) * (g->size() + 1));
  memcpy(colidx, g->edge_dst_host_ptr(), sizeof(uint32_t) * g->sizeEdges());
  if (vdata && g->vertex_data_) memcpy(vdata, g->vertex_data_, sizeof(float) * g->size());
  if (edata && g->edge_data_) memcpy(edata, g->edge_data_, sizeof(float) * g->sizeEdges());
}
void ref_graph_free(void* g) { ((Graph*)g)->dealloc(); delete (Graph*)g; }

void ref_set_threads(int n) { omp_set_num_threads(n); openblas_set_num_threads(n); }

void ref_init_glorot(size_t dx, size_t dy, float* out, unsigned seed) {
  vec_t w(dx * dy);
  init_glorot(dx, dy, w, seed);
  memcpy(out, w.data(), sizeof(float) * dx * dy);
}

// aggregators: reference classes called directly (include/gnn/aggregator.h:21-88)
void ref_gcn_aggregate(void* g, int len, const float* in, float* out) {
  GCN_Aggregator a; a.init(len, 0);
  a.aggregate(len, *(Graph*)g, in, out);
}
void ref_sage_aggregate(void* g, int len, const float* in, float* out, int transposed) {
  SAGE_Aggregator a; a.init(len, 0);
  if (transposed) a.d_aggregate(len, *(Graph*)g, NULL, in, out);
  else a.aggregate(len, *(Graph*)g, in, out);
}
// GAT aggregator forward+backward with caller-supplied attention vectors.
void ref_gat_aggregate(void* g_, int len, const float* alpha_l, const float* alpha_r, const float* z, float* out,
                       float* norm_scores, const float* grad_in, float* grad_out, float* dalpha_l, float* dalpha_r) {
  Graph* g = (Graph*)g_;
  GAT_Aggregator a; a.init(len, g->size(), g->sizeEdges(), 0.01f, 0.f);
  memcpy(a.alpha_l.data(), alpha_l, sizeof(float) * len);
  memcpy(a.alpha_r.data(), alpha_r, sizeof(float) * len);
  a.aggregate(len, *g, z, out);
  if (norm_scores) memcpy(norm_scores, a.norm_scores.data(), sizeof(float) * g->sizeEdges());
  if (grad_in) {
    a.d_aggregate(len, *g, z, grad_in, grad_out);
    memcpy(dalpha_l, a.alpha_lgrad.data(), sizeof(float) * len);
    memcpy(dalpha_r, a.alpha_rgrad.data(), sizeof(float) * len);
  }
}
void ref_symmetric_csr_transpose(int n, int nnz, const uint32_t* rowptr, const uint32_t* colidx, const float* vals, float* out) {
  float* t = NULL;
  symmetric_csr_transpose(n, nnz, (int*)rowptr, (int*)colidx, (float*)vals, t);
  memcpy(out, t, sizeof(float) * nnz);
  delete[] t;
}
void ref_matmul(size_t x, size_t y, size_t z, const float* A, const float* B, float* C, int ta, int tb, int accum) {
  matmul(x, y, z, A, B, C, ta, tb, accum);
}
// softmax loss on caller-supplied logits (src/layers/softmax_loss_layer.cpp:4-55, math_functions.cpp:79-92)
float ref_softmax_loss(int nv, int ncls, const float* logits, const uint8_t* labels, const uint8_t* masks, size_t begin,
                       size_t end, size_t count, float* probs, float* grad, float* acc) {
  softmax_loss_layer l(nv, ncls, (label_t*)labels);
  memcpy(l.get_feat_in(), logits, sizeof(float) * nv * ncls);
  l.forward(begin, end, (mask_t*)masks);
  float loss = l.get_prediction_loss(begin, end, count, (mask_t*)masks);
  if (probs) memcpy(probs, l.get_feat_out(), sizeof(float) * nv * ncls);
  if (grad) l.backward(begin, end, (mask_t*)masks, grad);
  if (acc) *acc = masked_accuracy_single(begin, end, count, ncls, (mask_t*)masks, l.get_feat_in(), (label_t*)labels);
  return loss;
}
// sigmoid (multi-label) loss on caller-supplied logits (src/layers/sigmoid_loss_layer.cpp:4-55) and the micro-F1 the reference
// reports as "accuracy" (masked_accuracy_multi -> masked_f1_score, math_functions.cpp:94-97,580-623). labels: [nv x ncls] multi-hot.
float ref_sigmoid_loss(int nv, int ncls, const float* logits, const uint8_t* labels, const uint8_t* masks, size_t begin, size_t end,
                       size_t count, float* probs, float* losses, float* grad, float* f1) {
  sigmoid_loss_layer l(nv, ncls, (label_t*)labels);
  memcpy(l.get_feat_in(), logits, sizeof(float) * nv * ncls);
  l.forward(begin, end, (mask_t*)masks);
  float loss = l.get_prediction_loss(begin, end, count, (mask_t*)masks);
  if (probs) memcpy(probs, l.get_feat_out(), sizeof(float) * nv * ncls);
  if (losses) for (size_t i = begin; i < end; i++) losses[i] = sigmoid_cross_entropy(ncls, (label_t*)labels + (size_t)ncls * i, logits + (size_t)ncls * i);
  if (grad) l.backward(begin, end, (mask_t*)masks, grad);
  if (f1) *f1 = masked_accuracy_multi(begin, end, count, ncls, (mask_t*)masks, l.get_feat_out(), (label_t*)labels);
  return loss;
}
// One Adam step with a fresh optimizer advanced `prior_calls` times (src/utilities/optimizer.cpp:22-35).
void ref_adam_steps(size_t n, float lr, int steps, const float* grads /*steps x n*/, float* w) {
  adam opt(lr);
  vec_t W(w, w + n), dW(n);
  for (int s = 0; s < steps; s++) {
    memcpy(dW.data(), grads + (size_t)s * n, sizeof(float) * n);
    opt.update(dW, W);
  }
  memcpy(w, W.data(), sizeof(float) * n);
}

// arch: 0 GCN, 1 SAGE, 2 GAT. `g` is the raw graph (no self-loops); the model takes ownership and adds them as net.cpp does.
void* ref_model_new(int arch, void* g, int nv, int dim_init, int dim_hid, int num_cls, int num_layers, float lr,
                    const float* feats, const uint8_t* labels, const int64_t* split9, int threads) {
  if (arch == 0) return new ModelBox<GCN_layer>((Graph*)g, nv, dim_init, dim_hid, num_cls, num_layers, lr, feats, labels, split9, threads);
  if (arch == 1) return new ModelBox<SAGE_layer>((Graph*)g, nv, dim_init, dim_hid, num_cls, num_layers, lr, feats, labels, split9, threads);
  if (arch == 2) return new ModelBox<GAT_layer>((Graph*)g, nv, dim_init, dim_hid, num_cls, num_layers, lr, feats, labels, split9, threads);
  return NULL;
}
// the same with the sigmoid (multi-label) loss layer and micro-F1 as accuracy (argv[4] == "sigmoid", net.cpp:20,447-451)
void* ref_model_new_sigmoid(int arch, void* g, int nv, int dim_init, int dim_hid, int num_cls, int num_layers, float lr,
                            const float* feats, const uint8_t* labels, const int64_t* split9, int threads) {
  if (arch == 0) return new ModelBox<GCN_layer>((Graph*)g, nv, dim_init, dim_hid, num_cls, num_layers, lr, feats, labels, split9, threads, true);
  if (arch == 1) return new ModelBox<SAGE_layer>((Graph*)g, nv, dim_init, dim_hid, num_cls, num_layers, lr, feats, labels, split9, threads, true);
  if (arch == 2) return new ModelBox<GAT_layer>((Graph*)g, nv, dim_init, dim_hid, num_cls, num_layers, lr, feats, labels, split9, threads, true);
  return NULL;
}
float ref_model_train_epoch(void* m, float* loss) { return ((ModelBase*)m)->train_epoch(loss); }
float ref_model_forward(void* m, float* loss) { return ((ModelBase*)m)->forward_only(loss); }
void ref_model_backward(void* m) { ((ModelBase*)m)->backward_only(); }
void ref_model_update(void* m) { ((ModelBase*)m)->update_only(); }
float ref_model_evaluate(void* m, const char* type) { return ((ModelBase*)m)->evaluate(type); }
int64_t ref_model_tensor_size(void* m, const char* name, int layer) {
  float* p = NULL;
  return ((ModelBase*)m)->tensor(name, layer, &p);
}
int64_t ref_model_get(void* m, const char* name, int layer, float* out, int64_t cap) {
  float* p = NULL;
  int64_t n = ((ModelBase*)m)->tensor(name, layer, &p);
  if (n < 0 || n > cap) return -1;
  memcpy(out, p, sizeof(float) * n);
  return n;
}
int64_t ref_model_set(void* m, const char* name, int layer, const float* in, int64_t n_in) {
  float* p = NULL;
  int64_t n = ((ModelBase*)m)->tensor(name, layer, &p);
  if (n < 0 || n != n_in) return -1;
  memcpy(p, in, sizeof(float) * n);
  return n;
}

// The reference's own loader (src/gnn/reader.cpp:248-457) on a dataset directory under DATASET_PATH (the path is read once, at
// library load: include/gnn/configs.h:5). Two-call protocol: with rowptr == NULL only the sizes come back in meta[0..12] =
// {nv, ne, feat_len, num_classes, train begin/end/count, val begin/end/count, test begin/end/count}.
int ref_reader_load(const char* dataset, int single_class, int64_t* meta, uint32_t* rowptr, uint32_t* colidx, float* feats, uint8_t* labels) {
  Reader reader{std::string(dataset)};
  Graph* g = new Graph(false);
  reader.bin_read_graph(g);
  std::vector<float> f;
  const size_t flen = reader.bin_read_features(f);
  std::vector<label_t> lab;
  const int ncls = reader.bin_read_vlabels(lab, single_class != 0);
  size_t b[3], e[3], c[3];
  const char* kinds[3] = {"train", "val", "test"};
  for (int i = 0; i < 3; i++) c[i] = reader.bin_read_masks(kinds[i], g->size(), b[i], e[i], NULL);
  meta[0] = g->size(); meta[1] = g->sizeEdges(); meta[2] = (int64_t)flen; meta[3] = ncls;
  for (int i = 0; i < 3; i++) { meta[4 + 3 * i] = b[i]; meta[5 + 3 * i] = e[i]; meta[6 + 3 * i] = c[i]; }
  if (rowptr) {
    memcpy(rowptr, g->row_start_host_ptr(), sizeof(uint32_t) * (g->size() + 1));
    memcpy(colidx, g->edge_dst_host_ptr(), sizeof(uint32_t) * g->sizeEdges());
    memcpy(feats, f.data(), sizeof(float) * f.size());
    memcpy(labels, lab.data(), lab.size());
  }
  return 0;
}

// the reference's legacy loader (src/gnn/reader.cpp:16-246) with the same protocol as ref_reader_load; masks_out = 3 x nv bytes
int ref_reader_load_csgr(const char* dataset, int single_class, int64_t* meta, uint32_t* rowptr, uint32_t* colidx, float* feats, uint8_t* labels,
                         uint8_t* masks_out) {
  Reader reader{std::string(dataset)};
  Graph* g = new Graph(false);
  reader.csgr_read_graph(g);
  std::vector<float> f;
  const size_t flen = reader.csgr_read_features(f, "bin");
  std::vector<label_t> lab;
  const size_t ncls = reader.csgr_read_labels(lab, single_class != 0);
  std::vector<mask_t> masks(3 * g->size(), 0);
  size_t b[3], e[3], c[3];
  const char* kinds[3] = {"train", "val", "test"};
  for (int i = 0; i < 3; i++) c[i] = reader.csgr_read_masks(kinds[i], g->size(), b[i], e[i], masks.data() + (size_t)i * g->size());
  meta[0] = g->size(); meta[1] = g->sizeEdges(); meta[2] = (int64_t)flen; meta[3] = (int64_t)ncls;
  for (int i = 0; i < 3; i++) { meta[4 + 3 * i] = b[i]; meta[5 + 3 * i] = e[i]; meta[6 + 3 * i] = c[i]; }
  if (rowptr) {
    memcpy(rowptr, g->row_start_host_ptr(), sizeof(uint32_t) * (g->size() + 1));
    memcpy(colidx, g->edge_dst_host_ptr(), sizeof(uint32_t) * g->sizeEdges());
    memcpy(feats, f.data(), sizeof(float) * f.size());
    memcpy(labels, lab.data(), lab.size());
    memcpy(masks_out, masks.data(), masks.size());
  }
  return 0;
}

// The reference's Sampler (src/gnn/sampler.cpp) on an in-memory graph: masked training graph (LearningGraph::generate_masked_graph,
// lgraph.h:231-272), frontier sampling (select_vertices, :170-294) and the induced, re-indexed subgraph (generateSubgraph, :148-158).
// Returns the number of selected vertices; set_out (ascending) / rowptr_out / colidx_out are filled when non-NULL; sizes = {n_set, nnz}.
int64_t ref_sampler_run(void* g_full, const uint8_t* masks_train, size_t count, uint32_t n, unsigned seed, int64_t* sizes, uint32_t* set_out,
                        uint32_t* rowptr_out, uint32_t* colidx_out) {
  Graph* full = (Graph*)g_full;
  Graph* tg = full->generate_masked_graph((mask_t*)masks_train);
  Sampler sampler(full, tg, (mask_t*)masks_train, count);
  VertexSet st;
  sampler.select_vertices(n, st, seed);
  std::vector<mask_t> masks(full->size(), 0);
  Graph sg;
  sampler.generateSubgraph(st, masks.data(), &sg);
  sizes[0] = (int64_t)st.size(); sizes[1] = (int64_t)sg.sizeEdges();
  if (set_out) {
    size_t i = 0;
    for (auto v : st) set_out[i++] = v;
    memcpy(rowptr_out, sg.row_start_host_ptr(), sizeof(uint32_t) * (sg.size() + 1));
    memcpy(colidx_out, sg.edge_dst_host_ptr(), sizeof(uint32_t) * sg.sizeEdges());
  }
  return (int64_t)st.size();
}

}  // extern "C"