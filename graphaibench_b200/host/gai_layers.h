// Host-side mirror of the reference's aggregator / layer / loss / optimizer classes over the C ABI.
//
//   optimizer, adam                      <- include/utils/optimizer.h:23-116, src/utilities/optimizer.{cpp,cu}
//   GCN_/SAGE_/GAT_Aggregator            <- include/gnn/aggregator.h:8-88, src/gnn/gconv/*_aggregator.{cpp,cu}
//   graph_conv_layer<A>, GCN_/SAGE_/GAT_layer <- include/layers/graph_conv_layer.h:6-106, src/gnn/gconv/*_layer.{cpp,cu}
//   loss_layer, softmax_loss_layer       <- include/gnn/loss_layer.h:5-31, src/layers/softmax_loss_layer.{cpp,cu}
//   dense_layer, l2norm_layer            <- include/layers/{dense,l2norm}_layer.h, src/layers/*.cpp
//
// All float* members and arguments are DEVICE pointers (as in the reference's GPU object set). Ownership follows the
// reference: a layer owns grad_in, its temporaries, weights and (level > 0) feat_in; forward(feat_out) writes into a
// buffer owned by the next layer. What differs is the schedule underneath: ReLU / "+self term" run as SpMM or GEMM
// epilogues, no per-launch device sync, no per-call allocations.
#pragma once
#include <iostream>
#include <unordered_map>
#include "gai_graph.h"

enum class net_phase { TRAIN, TEST, VAL };
enum class gnn_arch { GCN, GAT, SAGE, GGNN };

void init_glorot(size_t dim_x, size_t dim_y, vec_t& weight, unsigned seed);  // math_functions.cpp:11-19, host only
float* float_malloc_device_zero(size_t n);                                   // float_malloc_device + init_const_gpu(0)
void copy_float_to_device(size_t n, const float* src_h, float* dst_d);
void copy_float_to_host(size_t n, const float* src_d, float* dst_h);

// ---- optimizers ---------------------------------------------------------------------------------------------------
struct optimizer {
  virtual ~optimizer() {}
  virtual void update_gpu(const size_t n, const float* dW, float* W) = 0;
  virtual void reset() {}
};

// Adam with the reference's quirks: eps inside the sqrt, b1_t/b2_t advance once per update call on this object, first and
// second moments keyed by the address of W (optimizer.h:47-53, optimizer.cpp:22-35).
struct adam : public optimizer {
  explicit adam(float lr = 0.01f) : alpha(lr), b1(0.9f), b2(0.999f), b1_t(0.9f), b2_t(0.999f), eps(1e-8f) {}
  void update_gpu(const size_t n, const float* dW, float* W) override;
  void reset() override;
  float alpha, b1, b2, b1_t, b2_t;

 private:
  float eps;
  std::unordered_map<const float*, std::pair<float*, float*>> moments;
};

// Row pitch (floats) of every per-vertex activation / gradient buffer the layer classes own: rows padded to a multiple of 4 floats
// (47 classes -> pitch 48), so that every aggregation gather, TMA box and epilogue store is 16-byte aligned for any width.
// Measured and rejected (round 2, profiles/README.md): 128-byte-aligned rows (47 -> 64, 100 -> 128). They cut the L1 wavefronts of a
// gather by a third (a quarter-warp's 128 bytes no longer straddle two lines) but the aggregation is bound by L2 bandwidth and latency,
// not by the L1 data pipe, and the larger footprint lowered the L2 / L1 hit rates: F = 100 unchanged, F = 47 5-10 % slower.
inline size_t row_pitch(size_t dim) { return (dim + 3) / 4 * 4; }
inline size_t ceil4(size_t dim) { return (dim + 3) / 4 * 4; }
// Words per row of a sign-bit matrix (one bit per activation, GAI_EPI_BITMASK).
inline size_t bits_pitch(size_t dim) { return (dim + 31) / 32; }

// ---- aggregators --------------------------------------------------------------------------------------------------
// aggregate / d_aggregate keep the reference signatures (dense rows, aggregator.h:21-88); the *_ld forms take row pitches
// and epilogue flags (GAI_EPI_ADD with an addend of pitch ld_out, GAI_EPI_RELU) and are what the layers call.
class aggregator {
 public:
  void set_vlen(int vlen) { length = vlen; }

 protected:
  int length = 0;
};

// On a partitioned graph (Graph::partitioned()) `in` holds the master rows; the *_ld forms fetch the halo vertices' rows from their
// owners (Graph::halo_exchange) before aggregating the master rows, unless `static_halo` already holds them (layer-0 input features:
// fetched once by Model).
class GCN_Aggregator : public aggregator {
 public:
  void init(int len, int nv, int ne = 0, float lr = 0.01f, float drop_rate = 0.f);
  void aggregate(int len, Graph& g, const float* in, float* out);
  void d_aggregate(int len, Graph& g, const float* feat_in, const float* grad_in, float* grad_out);
  void aggregate_ld(int len, Graph& g, const float* in, size_t ld_in, float* out, size_t ld_out, int epilogue_flags, const float* addend,
                    const float* static_halo = nullptr);
  void d_aggregate_ld(int len, Graph& g, const float* grad_in, size_t ld_in, float* grad_out, size_t ld_out, int epilogue_flags, const float* addend,
                      const uint32_t* mask_bits = nullptr);
};

class SAGE_Aggregator : public aggregator {
 public:
  void init(int len, int nv, int ne = 0, float lr = 0.01f, float drop_rate = 0.f);
  void aggregate(int len, Graph& g, const float* in, float* out);
  void d_aggregate(int len, Graph& g, const float* feat_in, const float* grad_in, float* grad_out);
  void aggregate_ld(int len, Graph& g, const float* in, size_t ld_in, float* out, size_t ld_out, int epilogue_flags, const float* addend,
                    const float* static_halo = nullptr);
  void d_aggregate_ld(int len, Graph& g, const float* grad_in, size_t ld_in, float* grad_out, size_t ld_out, int epilogue_flags, const float* addend,
                      const uint32_t* mask_bits = nullptr);
};

class GAT_Aggregator : public aggregator {
 public:
  void init(int len, int nv, int ne = 0, float lr = 0.01f, float drop_rate = 0.f);
  void aggregate(int len, Graph& g, const float* in, float* out);
  void d_aggregate(int len, Graph& g, const float* feat_in, const float* grad_in, float* grad_out);
  void aggregate_fused(int len, Graph& g, const float* in, float* out, int epilogue_flags, const float* addend);
  void update_weights(optimizer* opt);
  float *d_alpha_l = nullptr, *d_alpha_r = nullptr, *d_alpha_lgrad = nullptr, *d_alpha_rgrad = nullptr;
  float *d_temp_scores = nullptr, *d_norm_scores = nullptr, *d_scores_grad = nullptr;

 private:
  float epsilon = 0.2f;  // LeakyReLU slope (gat_aggregator.cpp:22)
  float attn_drop = 0.f;
  int heads = 1;         // attention heads (GAI_GAT_HEADS; 1 = the reference, which has no multi-head path)
  optimizer* alpha_opt = nullptr;
 public:
  int num_heads() const { return heads; }
};

// ---- graph convolution layers -------------------------------------------------------------------------------------
template <typename Aggregator>
class graph_conv_layer {
 public:
  graph_conv_layer(int id, int nv, int din, int dout, Graph* g, bool act, bool concat, float lr, float feat_drop, float score_drop);
  float* get_feat_in() { return feat_in; }
  float* get_grad_in() { return grad_in; }
  void set_feat_in(float* ptr) { feat_in = ptr; }  // rows stored with row_pitch(dim_in), layer 0 included
  void set_graph_ptr(Graph* ptr) { graph = ptr; }
  void set_netphase(net_phase phase) { phase_ = phase; }
  void update_dim_size(size_t sz) { num_samples = (int)sz; }
  void print_layer_info() {
    gai_host::out() << "GraphConv Layer " << level_ << " with " << num_samples << " samples, dims: [" << dim_in << " x " << dim_out << "]\n";
  }
  // weight access for parity injection / checkpointing (host <-> device copies)
  int get_dim_in() const { return dim_in; }
  int get_dim_out() const { return dim_out; }
  float* weight_ptr(const std::string& name);  // "W", "W_grad", "W_self", "W_self_grad", "alpha_l", ... (device)
  size_t weight_size(const std::string& name);   // logical element count (rows x cols, dense)
  // Row layout of a named per-vertex tensor: logical columns and the pitch it is stored with (0/0 for weights: dense).
  void tensor_layout(const std::string& name, size_t* cols, size_t* ld);
  // Row pitches. Per-vertex activation / gradient buffers owned by the layer classes are stored with rows padded to 4 floats
  // (row_pitch: 47 classes -> pitch 48); layer 0's feat_in is Model's device copy of the input features, stored the same way.
  size_t ld_feat_in() const { return ld_in; }
  size_t ld_grad_in() const { return ld_out; }
  // d_relu fusion across the layer boundary: the layer above writes this layer's grad_in already masked by this layer's
  // activation (the GEMM epilogue reads its own feat_in = this layer's output), so backward() skips the separate pass.
  bool can_mask_grad_out() const;                 // this layer's last backward op on grad_out is a dense transform
  bool can_mask_grad_out_bits() const;            // ... or an aggregation, which can mask with the lower layer's sign bits only
  void set_mask_grad_out(bool on) { mask_grad_out = on; }
  void set_grad_premasked(bool on) { grad_premasked = on; }
  // Sign bits of this layer's activation (written by its forward when ReLU runs in a dense-transform epilogue; NULL otherwise) and
  // the bits of the layer below, which this layer's masked input gradient reads instead of the activation itself.
  const uint32_t* relu_bits() const { return d_relu_bits; }
  void set_mask_bits(const uint32_t* bits) { mask_bits_in = bits; }
  bool has_activation() const { return is_act; }
  // Partitioned graphs: matrices an aggregation gathers from are registered with the peer group; weight gradients are summed over the
  // ranks at the end of backward(). Layer 0's input never changes during training: Model fetches its halo rows once and hands them over.
  bool input_is_gathered() const { return !transform_first_; }
  void set_input_halo_static(const float* halo_rows) { input_static_halo = halo_rows; }

 protected:
  int level_, num_samples, dim_in, dim_out;
  Graph* graph;
  bool is_act, is_bias, use_concat;
  float feat_dropout_rate, score_dropout_rate, feat_scale;
  net_phase phase_ = net_phase::TRAIN;
  size_t ld_in = 0, ld_out = 0;  // row pitches of feat_in / (grad_in, out_temp, the feat_out this layer writes)
  bool mask_grad_out = false, grad_premasked = false;
  uint32_t* d_relu_bits = nullptr;
  const uint32_t* mask_bits_in = nullptr;
  bool transform_first_ = false;
  const float* input_static_halo = nullptr;
  float *d_W_neigh_grad_local = nullptr, *d_W_self_grad_local = nullptr;  // this rank's partial sums (== the public gradients when not partitioned)
  void reduce_weight_grads();
  // feature dropout (feat_dropout_rate > 0): the dropped-out input of the last training forward and its mask
  float* d_drop_in = nullptr;
  uint8_t* d_dropout_mask = nullptr;
  unsigned long long dropout_calls = 0;
  const float* forward_input();
  void backward_dropout(float* grad_out);
  float *feat_in = nullptr, *grad_in = nullptr;
  float *d_in_temp = nullptr, *d_in_temp1 = nullptr, *d_out_temp = nullptr;
  float *d_W_neigh = nullptr, *d_W_neigh_grad = nullptr, *d_W_self = nullptr, *d_W_self_grad = nullptr;
  optimizer* optm = nullptr;
  Aggregator aggr;
};

class GCN_layer : public graph_conv_layer<GCN_Aggregator> {
 public:
  GCN_layer(int id, int nv, int din, int dout, Graph* g, bool act, float lr, float feat_drop_rate, float score_drop_rate);
  void forward(float* feat_out);
  void backward(float* feat_out, float* grad_out);
  void update_weight(optimizer* opt);
};

class SAGE_layer : public graph_conv_layer<SAGE_Aggregator> {
 public:
  SAGE_layer(int id, int nv, int din, int dout, Graph* g, bool act, float lr, float feat_drop_rate, float score_drop_rate);
  void forward(float* feat_out);
  void backward(float* feat_out, float* grad_out);
  void update_weight(optimizer* opt);
};

class GAT_layer : public graph_conv_layer<GAT_Aggregator> {
 public:
  GAT_layer(int id, int nv, int din, int dout, Graph* g, bool act, float lr, float feat_drop_rate, float score_drop_rate);
  void forward(float* feat_out);
  void backward(float* feat_out, float* grad_out);
  void update_weight(optimizer* opt);
  GAT_Aggregator& aggregator_ref() { return aggr; }
};

// ---- tail layers ----------------------------------------------------------------------------------------------------
class l2norm_layer {
 public:
  l2norm_layer(int nv, int len);
  void forward(float* feat_out);
  void backward(float* grad_out);
  float* get_feat_in() { return feat_in; }
  float* get_grad_in() { return grad_in; }
  void update_dim_size(int sz) { num_samples = sz; }

 private:
  int num_samples, dim;
  float *feat_in, *grad_in;
};

class dense_layer {
 public:
  dense_layer(int nv, int in_len, int out_len, float lr);
  void forward(float* feat_out);
  void backward(float* grad_out);  // also applies its own Adam step (dense_layer.cpp:63-69)
  float* get_feat_in() { return feat_in; }
  float* get_grad_in() { return grad_in; }
  void update_dim_size(int sz) { num_samples = sz; }
  float *d_weight, *d_weight_grad;
  int dim_in, dim_out;

 private:
  int num_samples;
  float *feat_in, *grad_in;
  optimizer* optm;
};

class loss_layer {
 public:
  loss_layer(int nv, int ncls, label_t* ptr);
  virtual ~loss_layer() {}
  float* get_feat_in() { return feat_in; }
  float* get_feat_out() { return feat_out; }
  virtual void forward(size_t, size_t, mask_t*) {}
  virtual void backward(size_t, size_t, mask_t*, float*) {}
  virtual acc_t get_prediction_loss(size_t, size_t, size_t, mask_t*) { return 0; }
  void set_labels_ptr(label_t* ptr) { labels = ptr; }
  void set_netphase(net_phase phase) { phase_ = phase; }
  void update_dim_size(int sz) { num_samples = sz; }
  void print_layer_info() { gai_host::out() << "Output Layer with " << num_samples << " samples and " << num_cls << " classes\n"; }
  // 1D partition: this rank's rows are part of the reference's [begin, end) range. The gradient is scaled by the GLOBAL range length
  // (softmax_loss_layer.cpp:31) and the loss / accuracy statistics are combined over the ranks (collective: every rank calls).
  void set_partition(gai_host::Comm* comm);
  void set_global_denominator(size_t n) { global_denom = n; }

 protected:
  gai_host::Comm* comm_ = nullptr;
  size_t global_denom = 0;
  float* d_stats_all = nullptr;  // [world x 4] gathered statistics
  // {mean loss, accuracy, count} of this rank's rows in d_stats -> the same over all ranks' rows (host, double)
  void combine_stats(float* h3);
  int num_samples, num_cls;
  net_phase phase_ = net_phase::TRAIN;
  float *feat_in, *feat_out;
  label_t* labels;  // device
  float* d_losses;
  float* d_stats;   // {mean loss, accuracy, count}
};

class softmax_loss_layer : public loss_layer {
 public:
  softmax_loss_layer(int nv, int ncls, label_t* ptr) : loss_layer(nv, ncls, ptr) {}
  void forward(size_t begin, size_t end, mask_t* masks) override;
  void backward(size_t begin, size_t end, mask_t* masks, float* grad_out) override;
  acc_t get_prediction_loss(size_t begin, size_t end, size_t count, mask_t* masks) override;
  // masked_accuracy_single (math_functions.cpp:79-92) shares the reduction pass with the loss mean
  acc_t last_accuracy() const { return last_acc; }
  // accuracy of the current logits over [begin, end) — over every rank's rows when partitioned (Model::evaluate)
  acc_t masked_accuracy(size_t begin, size_t end, mask_t* masks);

 private:
  acc_t last_acc = 0;
  // rows the statistics in d_stats were computed over by the last forward()
  size_t stats_begin = 0, stats_end = 0;
  mask_t* stats_masks = nullptr;
  bool stats_valid = false;
};

// the output layer for multi-label classification (include/layers/sigmoid_loss_layer.h, src/layers/sigmoid_loss_layer.cpp:4-55);
// labels are [nv x ncls] multi-hot bytes on the device
class sigmoid_loss_layer : public loss_layer {
 public:
  sigmoid_loss_layer(int nv, int ncls, label_t* ptr) : loss_layer(nv, ncls, ptr) {}
  void forward(size_t begin, size_t end, mask_t* masks) override;
  void backward(size_t begin, size_t end, mask_t* masks, float* grad_out) override;
  acc_t get_prediction_loss(size_t begin, size_t end, size_t count, mask_t* masks) override;
};

float masked_accuracy_single(int begin, int end, int count, int num_classes, mask_t* masks, float* preds, label_t* ground_truth);
// micro-F1 at threshold 0.5 (masked_accuracy_multi -> masked_f1_score, math_functions.cpp:94-97,580-623); preds: the loss layer's
// feat_out (sigmoid outputs, rows stored with row_pitch), ground_truth: [nv x ncls] multi-hot
float masked_accuracy_multi(int begin, int end, int count, int num_classes, mask_t* masks, float* preds, label_t* ground_truth);
