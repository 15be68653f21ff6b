// Host-side mirror of the reference's model driver: Model<L> (include/gnn/net.h:9-83, src/gnn/net.cpp:11-620) over the
// layer classes in gai_layers.h. Same public methods and argv contract (train.cpp:9-15), subgraph sampling (argv subg_size > 0:
// Sampler + Model::subgraph_sampling, net.cpp:287-358) and inductive training (train on the training-masked graph, evaluate on the full
// one) included.
#pragma once
#include <string>
#include <vector>
#include "gai_dist.h"
#include "gai_layers.h"
#include "gai_sampler.h"

#define DEFAULT_NUM_LAYER 2
#define DEFAULT_SIZE_HID 16
#define DEFAULT_RATE_LEARN 0.02
#define EVAL_INTERVAL 50

template <typename gconv_layer>
class Model {
 public:
  Model() {}
  acc_t forward_prop(acc_t& loss);
  acc_t evaluate(std::string type);
  void backward_prop();
  void load_data(int argc, char* argv[]);
  void construct_network();
  void train();
  void update_weights(optimizer* opt);
  void set_netphases(net_phase phase);
  void print_layers_info();
  void transfer_data_to_device();

  // In-memory set-up: what load_data does after the Reader calls (net.cpp:88-203), for callers that already hold the
  // arrays (tests, bench.py through gai_host_capi.cpp). `split9` = train/val/test (begin, end, count).
  void init_from_memory(gnn_arch arch, Graph* g, int dim_init, int num_cls, const float* feats_h, const label_t* labels_h,
                        const int64_t* split9, int dim_hid, int num_layers, float lr, int epochs, int val_interval);
  // 1D-partitioned training (host/gai_dist.h): call set_comm() before load_data() / init_partitioned(). Every rank builds
  // the same network over its own masters; aggregations exchange halo rows, weight gradients and loss statistics are combined over the
  // ranks (csrc/peers.cu), weights and optimiser state are replicated and stay bit-identical on every rank.
  void set_comm(gai_host::Comm* c) { comm_ = c; gai_host::set_quiet(c && c->rank() != 0); }
  // This rank's rows of the RAW graph (global column ids, no self-loops; rows_rowptr has n_loc + 1 entries, rebased to 0), its rows of
  // the feature matrix and label vector, and the GLOBAL split ranges.
  void init_partitioned(gnn_arch arch, index_t nv_global, const int64_t* rows_rowptr, const uint32_t* rows_colidx, int dim_init, int num_cls,
                        const float* feats_local, const label_t* labels_local, const int64_t* split9_global, int dim_hid, int num_layers, float lr);
  Graph* graph() { return training_graph; }
  // One training step exactly as Model::train's loop body (net.cpp:373-383); returns train accuracy.
  acc_t train_epoch(acc_t& loss);
  // Re-upload the epoch's inputs (features, labels, masks, CSR) from host memory: the reference's host->device boundary
  // (net.cpp:186-187, 207-227), exposed so that an end-to-end step can include it. `feats_h` = dense [nv x dim_init] rows, read from
  // where they lie on EVERY call (pass page-locked memory for a true async copy); NULL = the Model's own copy of the training features.
  void refresh_inputs_from_host(const float* feats_h);
  // Start copying the NEXT step's feature matrix (the bulk of the inputs) into a second device buffer on a copy stream, behind
  // everything already enqueued on the compute stream; the following refresh_inputs_from_host() swaps it in instead of
  // copying in line, so the transfer overlaps the current step's kernels.
  void prefetch_features_from_host(const float* feats_h);
  int num_conv_layers() const { return num_layers; }
  gconv_layer& conv_layer(int l) { return layer_gconv[l]; }
  dense_layer* dense() { return layer_dense; }
  loss_layer* loss() { return layer_loss; }
  l2norm_layer* l2norm() { return layer_l2norm; }
  optimizer* shared_optimizer() { return opt_; }
  size_t nv() const { return num_samples; }

 private:
  int num_epochs = 0, num_layers = DEFAULT_NUM_LAYER, num_samples = 0, num_threads = 1, num_cls = 0, dim_init = 0;
  int dim_hid = DEFAULT_SIZE_HID, subg_size = 0, val_interval = EVAL_INTERVAL;
  float feat_drop = 0.f, score_drop = 0.f, lrate = DEFAULT_RATE_LEARN;
  size_t train_begin = 0, train_end = 0, train_count = 0, val_begin = 0, val_end = 0, val_count = 0, test_begin = 0, test_end = 0, test_count = 0;
  Graph* full_graph = nullptr;
  Graph* training_graph = nullptr;
  std::string dataset_name;
  bool is_sigmoid = false, use_dense = false, use_l2norm = false, use_gpu = true, inductive = false;
  gnn_arch arch = gnn_arch::GCN;
  std::vector<gconv_layer> layer_gconv;
  loss_layer* layer_loss = nullptr;
  l2norm_layer* layer_l2norm = nullptr;
  dense_layer* layer_dense = nullptr;
  optimizer* opt_ = nullptr;
  std::vector<float> input_features;
  std::vector<label_t> labels;
  std::vector<mask_t> masks_train, masks_test, masks_val;
  float* d_input_features = nullptr;
  float* d_input_halo = nullptr;  // partitioned: feature rows of the halo vertices (fetched once)
  float* d_feat_buf[2] = {nullptr, nullptr};  // double-buffered input features (prefetch_features_from_host)
  int feat_cur = 0;
  bool prefetch_pending = false;
  void* copy_stream = nullptr;
  void *ev_ready = nullptr, *ev_free = nullptr;
  void* pinned_inputs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  void stage_pinned();
  const float* feature_source(const float* feats_h);
  void upload_features(float* dst_d, const float* src_h, void* on_stream);
  label_t* d_labels = nullptr;
  mask_t *d_masks_train = nullptr, *d_masks_test = nullptr, *d_masks_val = nullptr;

  gai_host::Comm* comm_ = nullptr;
  bool partitioned() const { return comm_ != nullptr; }
  bool talk() const { return !comm_ || comm_->rank() == 0; }  // only rank 0 logs
  size_t train_denominator = 0;   // GLOBAL length of the train range (the reference's end - begin, softmax_loss_layer.cpp:31)
  void localise_split(const int64_t* split9_global, index_t first, index_t last);
  void finish_setup();
  void run_forward_layers();
  // subgraph sampling (net.cpp:64-79,154-186,287-358)
  int epochs_done = 0;
  Sampler* sampler = nullptr;
  int num_subgraphs = 1, subg_nv = 0, num_subg_remain = 0;
  std::vector<Graph*> subgs;
  std::vector<mask_t> subg_masks;   // num_subgraphs x num_samples
  float* d_feats_subg = nullptr;
  label_t* d_labels_subg = nullptr;
  uint32_t* d_subg_ids = nullptr;   // kept vertex ids of the subgraph in use (device), for the feature-row gather
  std::vector<label_t> labels_subg;
  void subgraph_sampling(int cur_epoch);
  void use_full_graph();            // evaluate(): back to the full graph, features and labels
};
