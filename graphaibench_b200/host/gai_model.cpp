#include "gai_model.h"
#include <cassert>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iomanip>

using gai_host::die_on;
using gai_host::stream;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void sync() { die_on(gai_stream_sync(stream()), "gai_stream_sync"); }

template <typename T>
static T* upload(const T* src, size_t n) {
  void* p = nullptr;
  die_on(gai_malloc(&p, sizeof(T) * (n ? n : 1)), "gai_malloc");
  die_on(gai_memcpy_h2d(p, src, sizeof(T) * n, stream()), "gai_memcpy_h2d");
  return reinterpret_cast<T*>(p);
}

template <typename L> struct arch_of;
template <> struct arch_of<GCN_layer> { static constexpr gnn_arch value = gnn_arch::GCN; };
template <> struct arch_of<SAGE_layer> { static constexpr gnn_arch value = gnn_arch::SAGE; };
template <> struct arch_of<GAT_layer> { static constexpr gnn_arch value = gnn_arch::GAT; };

template <typename L>
void Model<L>::load_data(int argc, char* argv[]) {
  // positional arguments as the reference (net.cpp:13-64)
  dataset_name = argv[1];
  num_epochs = atoi(argv[2]);
  num_threads = atoi(argv[3]);  // accepted for CLI parity; the device path has no host thread fan-out
  is_sigmoid = std::string(argv[4]) == "sigmoid";
  arch = arch_of<L>::value;
  if (argc >= 6) dim_hid = atoi(argv[5]);
  if (argc >= 7) score_drop = (float)atof(argv[6]);
  if (argc >= 8) feat_drop = (float)atof(argv[7]);
  if (argc >= 9) lrate = (float)atof(argv[8]);
  if (argc > 9) {
    assert(argc == 13);
    num_layers = atoi(argv[9]);
    subg_size = atoi(argv[10]);
    val_interval = atoi(argv[11]);
    inductive = atoi(argv[12]) != 0;
  }
  assert(num_layers >= 2);
  if (subg_size > 0) inductive = true;  // net.cpp:157
  full_graph = new Graph(true);
  Reader reader(dataset_name);
  reader.bin_read_graph(full_graph);
  num_samples = (int)full_graph->size();
  dim_init = (int)reader.bin_read_features(input_features);
  num_cls = reader.bin_read_vlabels(labels, !is_sigmoid);
  if (arch != gnn_arch::SAGE) full_graph->add_selfloop();  // net.cpp:96
  gai_host::out() << "num_threads = " << num_threads << ", num_vertices = " << num_samples << ", num_edges = " << full_graph->sizeEdges()
            << ", num_layers = " << num_layers << ", \nnum_epochs = " << num_epochs << ", input_length = " << dim_init
            << ", hidden_length = " << dim_hid << ", num_classes = " << num_cls << ", \nfeat_drop = " << feat_drop
            << ", score_drop = " << score_drop << ", subg_size = " << subg_size << ", val_interval = " << val_interval
            << ", learning_rate = " << lrate << "\n";
  train_count = reader.bin_read_masks("train", num_samples, train_begin, train_end, nullptr);
  val_count = reader.bin_read_masks("val", num_samples, val_begin, val_end, nullptr);
  test_count = reader.bin_read_masks("test", num_samples, test_begin, test_end, nullptr);
  if (partitioned() && (subg_size > 0 || inductive)) { std::cerr << "partitioned training runs the full graph only (no sampling / inductive mode)\n"; std::exit(1); }
  if (subg_size > 0 && val_interval < num_epochs) {  // net.cpp:100-103
    gai_host::out() << "disabling validation for subgraph sampling on GPU\n";
    val_interval = num_epochs;
  }
  if (partitioned()) {
    // every rank read the whole dataset; keep this rank's rows (global column ids, self-loops already in place), features and labels
    if (is_sigmoid) { std::cerr << "partitioned training supports the softmax loss only\n"; std::exit(1); }
    const index_t nv_global = (index_t)num_samples;
    const gai_host::OwnerRange own = gai_host::owner_range(nv_global, comm_->world(), comm_->rank());
    Graph* mine = new Graph(true);
    const index_t* rp = full_graph->row_start_host_ptr();
    const index_t* ci = full_graph->edge_dst_host_ptr();
    mine->allocateFrom(own.last - own.first, rp[own.last] - rp[own.first]);
    for (index_t v = own.first; v < own.last; v++) mine->fixEndEdge(v - own.first, rp[v + 1] - rp[own.first]);
    std::copy(ci + rp[own.first], ci + rp[own.last], mine->edge_dst_host_ptr());
    delete full_graph;
    full_graph = mine;
    const int64_t split9[9] = {(int64_t)train_begin, (int64_t)train_end, (int64_t)train_count, (int64_t)val_begin, (int64_t)val_end, (int64_t)val_count,
                               (int64_t)test_begin, (int64_t)test_end, (int64_t)test_count};
    std::vector<float> f(input_features.begin() + (size_t)own.first * dim_init, input_features.begin() + (size_t)own.last * dim_init);
    input_features.swap(f);
    std::vector<label_t> l(labels.begin() + own.first, labels.begin() + own.last);
    labels.swap(l);
    num_samples = (int)(own.last - own.first);
    localise_split(split9, own.first, own.last);
    full_graph->partition_rows(comm_, nv_global);
  }
  finish_setup();
}

template <typename L>
void Model<L>::localise_split(const int64_t* s, index_t first, index_t last) {
  // the reference's ranges are global row ranges; this rank keeps their intersection with its masters, as local rows
  auto clip = [&](int64_t v) { return (size_t)(std::min<int64_t>(std::max<int64_t>(v, first), last) - first); };
  train_denominator = (size_t)(s[1] - s[0]);
  train_begin = clip(s[0]); train_end = clip(s[1]); train_count = train_end - train_begin;
  val_begin = clip(s[3]); val_end = clip(s[4]); val_count = val_end - val_begin;
  test_begin = clip(s[6]); test_end = clip(s[7]); test_count = test_end - test_begin;
}

template <typename L>
void Model<L>::init_partitioned(gnn_arch a, index_t nv_global, const int64_t* rows_rowptr, const uint32_t* rows_colidx, int dinit, int ncls,
                                const float* feats_local, const label_t* labels_local, const int64_t* split9, int dhid, int nlayers, float lr) {
  assert(a == arch_of<L>::value && comm_ != nullptr);
  const gai_host::OwnerRange own = gai_host::owner_range(nv_global, comm_->world(), comm_->rank());
  const index_t n_loc = own.last - own.first;
  arch = a; num_samples = (int)n_loc; dim_init = dinit; num_cls = ncls; dim_hid = dhid; num_layers = nlayers; lrate = lr;
  num_epochs = 0; val_interval = 1 << 30;
  full_graph = new Graph(true);
  full_graph->allocateFrom(n_loc, (index_t)rows_rowptr[n_loc]);
  for (index_t v = 0; v < n_loc; v++) full_graph->fixEndEdge(v, (index_t)rows_rowptr[v + 1]);
  std::copy(rows_colidx, rows_colidx + rows_rowptr[n_loc], full_graph->edge_dst_host_ptr());
  input_features.assign(feats_local, feats_local + (size_t)n_loc * dinit);
  labels.assign(labels_local, labels_local + n_loc);
  localise_split(split9, own.first, own.last);
  if (arch != gnn_arch::SAGE) full_graph->add_selfloop_rows(own.first);
  full_graph->partition_rows(comm_, nv_global);
  finish_setup();
}

template <typename L>
void Model<L>::init_from_memory(gnn_arch a, Graph* g, int dinit, int ncls, const float* feats_h, const label_t* labels_h, const int64_t* s,
                                int dhid, int nlayers, float lr, int epochs, int vint) {
  assert(a == arch_of<L>::value);
  arch = a; full_graph = g; num_samples = (int)g->size(); dim_init = dinit; num_cls = ncls; dim_hid = dhid; num_layers = nlayers;
  lrate = lr; num_epochs = epochs; val_interval = vint;
  input_features.assign(feats_h, feats_h + (size_t)num_samples * dinit);
  labels.assign(labels_h, labels_h + num_samples);
  train_begin = s[0]; train_end = s[1]; train_count = s[2];
  val_begin = s[3]; val_end = s[4]; val_count = s[5];
  test_begin = s[6]; test_end = s[7]; test_count = s[8];
  if (arch != gnn_arch::SAGE) full_graph->add_selfloop();
  finish_setup();
}

template <typename L>
void Model<L>::finish_setup() {
  // l2norm + dense tail for GAT (net.cpp:67-71); masks from the meta ranges (net.cpp:126-142)
  use_l2norm = (arch == gnn_arch::GAT) || subg_size > 0;  // "l2norm+dense layer is useful for sampling and GAT" (net.cpp:67-71)
  use_dense = use_l2norm;
  masks_train.assign(num_samples, 0); masks_val.assign(num_samples, 0); masks_test.assign(num_samples, 0);
  for (size_t i = train_begin; i < train_end; i++) masks_train[i] = 1;
  for (size_t i = val_begin; i < val_end; i++) masks_val[i] = 1;
  for (size_t i = test_begin; i < test_end; i++) masks_test[i] = 1;
  training_graph = full_graph;
  if (inductive) {
    // net.cpp:154-172: train on the graph masked to the training vertices, evaluate on the full one; with sampling the masked graph is
    // what the frontier walk reads and the subgraphs are induced from the full graph
    assert((size_t)subg_size <= train_count);
    training_graph = full_graph->generate_masked_graph(masks_train.data());
    transfer_data_to_device();
    full_graph->copy_to_gpu();
    if (subg_size > 0) {
      num_subgraphs = num_threads > 0 ? num_threads : 1;
      sampler = new Sampler(full_graph, training_graph, masks_train.data(), train_count);
      subgs.assign(num_subgraphs, nullptr);
      subg_masks.assign((size_t)num_subgraphs * num_samples, 0);
      d_feats_subg = float_malloc_device_zero((size_t)subg_size * row_pitch(dim_init));
      void* p = nullptr;
      die_on(gai_malloc(&p, (size_t)subg_size * (is_sigmoid ? num_cls : 1)), "gai_malloc"); d_labels_subg = reinterpret_cast<label_t*>(p);
      die_on(gai_malloc(&p, sizeof(uint32_t) * (size_t)subg_size), "gai_malloc"); d_subg_ids = reinterpret_cast<uint32_t*>(p);
    } else {
      training_graph->copy_to_gpu();
    }
    return;
  }
  if (partitioned()) {  // the device graph (and its halo plan) first: the input features carry a halo block that is fetched through it
    training_graph->copy_to_gpu();
    transfer_data_to_device();
    return;
  }
  transfer_data_to_device();
  training_graph->alloc_on_device();
  training_graph->copy_to_gpu();
  training_graph->compute_edge_data();
}

template <typename L>
void Model<L>::transfer_data_to_device() {  // net.cpp:207-227
  // input features live on the device with rows of pitch row_pitch(dim_init), like every other per-vertex buffer
  d_input_features = float_malloc_device_zero((size_t)num_samples * row_pitch(dim_init));
  upload_features(d_input_features, input_features.data(), stream());
  if (partitioned()) {
    // the input never changes during training: the halo vertices' feature rows are fetched once, into a buffer of their own
    training_graph->register_gather_buffer(d_input_features);
    d_input_halo = float_malloc_device_zero(std::max<size_t>(training_graph->num_halo(), 1) * row_pitch(dim_init));
    training_graph->halo_exchange_into(d_input_features, dim_init, row_pitch(dim_init), d_input_halo);
  }
  d_labels = upload(labels.data(), labels.size());
  d_masks_train = upload(masks_train.data(), masks_train.size());
  d_masks_test = upload(masks_test.data(), masks_test.size());
  d_masks_val = upload(masks_val.data(), masks_val.size());
  sync();
}

// dense host rows [num_samples x dim_init] -> device rows of pitch row_pitch(dim_init): one linear DMA when the width is a multiple of
// 4 floats (every BASELINE.json shape), a pitched one otherwise (measured: a pitched host->device copy of 400-byte rows runs at
// ~4 GB/s, so wide matrices whose width is not a multiple of 4 should be padded on the host)
template <typename L>
void Model<L>::upload_features(float* dst_d, const float* src_h, void* on_stream) {
  const size_t w = sizeof(float) * (size_t)dim_init;
  if (row_pitch(dim_init) == (size_t)dim_init) die_on(gai_memcpy_h2d(dst_d, src_h, w * (size_t)num_samples, on_stream), "gai_memcpy_h2d");
  else die_on(gai_memcpy2d(dst_d, sizeof(float) * row_pitch(dim_init), src_h, w, w, (size_t)num_samples, on_stream), "gai_memcpy2d");
}

template <typename L>
void Model<L>::stage_pinned() {
  // The four inputs the Model itself owns (labels, train mask, CSR) and its own copy of the features are staged once in page-locked
  // host memory so that the per-step copies are true async DMA. They cannot change behind the Model's back; a feature matrix the
  // CALLER passes to refresh_inputs_from_host / prefetch_features_from_host is never staged: it is copied from where it lies, every call.
  if (pinned_inputs[1]) return;
  const size_t bytes[5] = {sizeof(float) * input_features.size(), labels.size(), masks_train.size(),
                           sizeof(index_t) * (training_graph->size() + 1), sizeof(index_t) * training_graph->sizeEdges()};
  const void* srcs[5] = {input_features.data(), labels.data(), masks_train.data(), training_graph->row_start_host_ptr(),
                         training_graph->edge_dst_host_ptr()};
  for (int i = 1; i < 5; i++) {
    die_on(gai_host_alloc_pinned(&pinned_inputs[i], bytes[i]), "gai_host_alloc_pinned");
    memcpy(pinned_inputs[i], srcs[i], bytes[i]);
  }
}

// The feature matrix a host->device step copies: the caller's (page-locked for a true async copy; read on every call, so new data
// is honoured) or, with NULL, the Model's own copy of the training features, pinned on first use.
template <typename L>
const float* Model<L>::feature_source(const float* feats_h) {
  if (feats_h) return feats_h;
  if (!pinned_inputs[0]) {
    die_on(gai_host_alloc_pinned(&pinned_inputs[0], sizeof(float) * input_features.size()), "gai_host_alloc_pinned");
    memcpy(pinned_inputs[0], input_features.data(), sizeof(float) * input_features.size());
  }
  return reinterpret_cast<const float*>(pinned_inputs[0]);
}

template <typename L>
void Model<L>::refresh_inputs_from_host(const float* feats_h) {
  stage_pinned();
  const size_t bytes[5] = {0, labels.size(), masks_train.size(), sizeof(index_t) * (training_graph->size() + 1),
                           sizeof(index_t) * training_graph->sizeEdges()};
  void* dsts[5] = {nullptr, d_labels, d_masks_train, (void*)training_graph->row_start_ptr(), (void*)training_graph->edge_dst_ptr()};
  for (int i = 1; i < 5; i++) die_on(gai_memcpy_h2d(dsts[i], pinned_inputs[i], bytes[i], stream()), "gai_memcpy_h2d");
  if (prefetch_pending) {  // the features of this step were sent ahead: order the compute stream behind the copy and swap buffers
    die_on(gai_stream_wait_event(stream(), ev_ready), "gai_stream_wait_event");
    feat_cur ^= 1;
    d_input_features = d_feat_buf[feat_cur];
    layer_gconv[0].set_feat_in(d_input_features);
    prefetch_pending = false;
  } else {
    upload_features(d_input_features, feature_source(feats_h), stream());
  }
  if (partitioned()) training_graph->halo_exchange_into(d_input_features, dim_init, row_pitch(dim_init), d_input_halo);  // new features: new halo rows
}

template <typename L>
void Model<L>::prefetch_features_from_host(const float* feats_h) {
  if (partitioned()) return;  // the double-buffered copy stream is a single-GPU path: partitioned steps copy in line (refresh_inputs_from_host)
  if (!copy_stream) {
    die_on(gai_stream_create(&copy_stream), "gai_stream_create");
    die_on(gai_event_create(&ev_ready), "gai_event_create");
    die_on(gai_event_create(&ev_free), "gai_event_create");
    d_feat_buf[0] = d_input_features;
    d_feat_buf[1] = float_malloc_device_zero((size_t)num_samples * row_pitch(dim_init));
  }
  if (prefetch_pending) return;  // one step ahead at most
  // the spare buffer was last read by the step before the current one: everything enqueued so far must finish first
  die_on(gai_event_record(ev_free, stream()), "gai_event_record");
  die_on(gai_stream_wait_event(copy_stream, ev_free), "gai_stream_wait_event");
  upload_features(d_feat_buf[feat_cur ^ 1], feature_source(feats_h), copy_stream);
  die_on(gai_event_record(ev_ready, copy_stream), "gai_event_record");
  prefetch_pending = true;
}

template <typename L>
void Model<L>::construct_network() {  // net.cpp:422-453
  gai_host::out() << "constructing neural network...\n";
  const int nv = num_samples;
  layer_gconv.reserve(num_layers);
  // inductive / sampling runs evaluate on the full graph: per-edge buffers (GAT) are sized by it; the layers then point at the training graph
  Graph* sizing_graph = inductive ? full_graph : training_graph;
  for (int l = 0; l < num_layers - 1; l++)
    layer_gconv.emplace_back(l, nv, l == 0 ? dim_init : dim_hid, dim_hid, sizing_graph, true, lrate, feat_drop, score_drop);
  layer_gconv.emplace_back(num_layers - 1, nv, dim_hid, use_dense ? dim_hid : num_cls, sizing_graph, false, lrate, feat_drop, score_drop);
  if (inductive) for (auto& y : layer_gconv) y.set_graph_ptr(training_graph);
  if (use_l2norm) layer_l2norm = new l2norm_layer(nv, dim_hid);
  if (use_dense) layer_dense = new dense_layer(nv, dim_hid, num_cls, lrate);
  layer_gconv[0].set_feat_in(d_input_features);
  if (partitioned()) layer_gconv[0].set_input_halo_static(d_input_halo);
  // d_relu of layer l-1 (gcn_layer.cpp:38-40) rides the epilogue of the transform that produces its grad_in in layer l
  for (int l = 1; l < num_layers; l++) {
    if (layer_gconv[l - 1].has_activation() &&
        (layer_gconv[l].can_mask_grad_out() || (layer_gconv[l].can_mask_grad_out_bits() && layer_gconv[l - 1].relu_bits()))) {
      layer_gconv[l].set_mask_grad_out(true);
      layer_gconv[l].set_mask_bits(layer_gconv[l - 1].relu_bits());  // NULL: mask with the activation itself
      layer_gconv[l - 1].set_grad_premasked(true);
    }
  }
  if (is_sigmoid) layer_loss = new sigmoid_loss_layer(nv, num_cls, d_labels);  // net.cpp:447-451
  else layer_loss = new softmax_loss_layer(nv, num_cls, d_labels);
  if (partitioned()) {
    layer_loss->set_partition(comm_);
    layer_loss->set_global_denominator(train_denominator);
  }
  opt_ = new adam(lrate);  // net.cpp:362
  sync();
}

template <typename L>
void Model<L>::update_weights(optimizer* opt) { for (int i = 0; i < num_layers; i++) layer_gconv[i].update_weight(opt); }
template <typename L>
void Model<L>::set_netphases(net_phase phase) { for (auto& y : layer_gconv) y.set_netphase(phase); layer_loss->set_netphase(phase); }
template <typename L>
void Model<L>::print_layers_info() { for (auto& y : layer_gconv) y.print_layer_info(); layer_loss->print_layer_info(); }

template <typename L>
void Model<L>::run_forward_layers() {  // shared by forward_prop and evaluate (net.cpp:458-471, 541-554)
  for (int l = 0; l < num_layers - 1; l++) layer_gconv[l].forward(layer_gconv[l + 1].get_feat_in());
  if (use_dense) {
    layer_gconv[num_layers - 1].forward(layer_l2norm->get_feat_in());
    layer_l2norm->forward(layer_dense->get_feat_in());
    layer_dense->forward(layer_loss->get_feat_in());
  } else {
    layer_gconv[num_layers - 1].forward(layer_loss->get_feat_in());
  }
}

template <typename L>
acc_t Model<L>::forward_prop(acc_t& loss) {
  run_forward_layers();
  size_t begin = train_begin, end = train_end, count = train_count;
  mask_t* masks_ptr = d_masks_train;
  label_t* labels_ptr = d_labels;
  if (subg_size > 0) { masks_ptr = nullptr; begin = 0; end = (size_t)subg_nv; count = (size_t)subg_nv; labels_ptr = d_labels_subg; }  // net.cpp:477-488
  layer_loss->forward(begin, end, masks_ptr);
  loss = layer_loss->get_prediction_loss(begin, end, count, masks_ptr);
  if (is_sigmoid)  // net.cpp:495-497
    return masked_accuracy_multi((int)begin, (int)end, (int)count, num_cls, masks_ptr, layer_loss->get_feat_out(), labels_ptr);
  return static_cast<softmax_loss_layer*>(layer_loss)->last_accuracy();  // same reduction pass as the loss mean
}

template <typename L>
acc_t Model<L>::evaluate(std::string type) {
  set_netphases(net_phase::TEST);
  if (subg_size > 0 || inductive) use_full_graph();  // net.cpp:508-535
  run_forward_layers();
  if (is_sigmoid) {  // net.cpp:569-572: the loss layer's forward produces the sigmoid outputs the F1 score thresholds
    const bool test = type == "test";
    const size_t b = test ? test_begin : val_begin, e = test ? test_end : val_end, c = test ? test_count : val_count;
    mask_t* mk = test ? d_masks_test : d_masks_val;
    layer_loss->forward(b, e, mk);
    return masked_accuracy_multi((int)b, (int)e, (int)c, num_cls, mk, layer_loss->get_feat_out(), d_labels);
  }
  if (partitioned()) {
    const bool test = type == "test";
    return static_cast<softmax_loss_layer*>(layer_loss)->masked_accuracy(test ? test_begin : val_begin, test ? test_end : val_end,
                                                                         test ? d_masks_test : d_masks_val);
  }
  if (type == "test") return masked_accuracy_single((int)test_begin, (int)test_end, (int)test_count, num_cls, d_masks_test, layer_loss->get_feat_in(), d_labels);
  return masked_accuracy_single((int)val_begin, (int)val_end, (int)val_count, num_cls, d_masks_val, layer_loss->get_feat_in(), d_labels);
}

template <typename L>
void Model<L>::backward_prop() {  // net.cpp:580-615
  size_t train_begin = this->train_begin, train_end = this->train_end;
  mask_t* d_masks_train = this->d_masks_train;
  if (subg_size > 0) { d_masks_train = nullptr; train_begin = 0; train_end = (size_t)subg_nv; }  // net.cpp:586-590
  if (use_dense) {
    layer_loss->backward(train_begin, train_end, d_masks_train, layer_dense->get_grad_in());
    layer_dense->backward(layer_l2norm->get_grad_in());
    layer_l2norm->backward(layer_gconv[num_layers - 1].get_grad_in());
    layer_gconv[num_layers - 1].backward(layer_l2norm->get_feat_in(), layer_gconv[num_layers - 2].get_grad_in());
  } else {
    layer_loss->backward(train_begin, train_end, d_masks_train, layer_gconv[num_layers - 1].get_grad_in());
    layer_gconv[num_layers - 1].backward(layer_loss->get_feat_in(), layer_gconv[num_layers - 2].get_grad_in());
  }
  for (int l = num_layers - 2; l > 0; l--) layer_gconv[l].backward(layer_gconv[l + 1].get_feat_in(), layer_gconv[l - 1].get_grad_in());
  layer_gconv[0].backward(layer_gconv[1].get_feat_in(), nullptr);
}

template <typename L>
acc_t Model<L>::train_epoch(acc_t& loss) {
  if (subg_size > 0) subgraph_sampling(epochs_done);
  else if (inductive) for (auto& y : layer_gconv) y.set_graph_ptr(training_graph);  // back from an evaluation on the full graph
  epochs_done++;
  set_netphases(net_phase::TRAIN);
  acc_t acc = forward_prop(loss);
  backward_prop();
  update_weights(opt_);
  return acc;
}

template <typename L>
void Model<L>::train() {  // log lines as net.cpp:364-410 so that logs diff cleanly against the reference
  gai_host::out() << "Start training...\n";
  double total_train_time = 0.0;
  for (int itr = 0; itr < num_epochs; itr++) {
    if (subg_size > 0) subgraph_sampling(itr);
    else if (inductive) for (auto& y : layer_gconv) y.set_graph_ptr(training_graph);
    gai_host::out() << "Epoch " << std::setw(3) << itr << " ";
    set_netphases(net_phase::TRAIN);
    acc_t train_loss = 0.0;
    const double t0 = now_s();
    acc_t train_acc = forward_prop(train_loss);
    const double t1 = now_s();
    backward_prop();
    update_weights(opt_);
    sync();
    const double t2 = now_s();
    const double fw_time = t1 - t0, bw_time = t2 - t1, epoch_time = fw_time + bw_time;
    total_train_time += epoch_time;
    gai_host::out() << "train_loss " << std::setprecision(3) << std::fixed << train_loss << " train_acc " << train_acc << " ";
    if (itr % val_interval == 0 && itr != 0) {
      const double v0 = now_s();
      acc_t val_acc = evaluate("val");
      const double val_time = now_s() - v0;
      gai_host::out() << "val_acc " << std::setprecision(3) << std::fixed << val_acc << " ";
      gai_host::out() << "time " << std::setprecision(3) << std::fixed << epoch_time + val_time << " s (train_time " << epoch_time << " val_time "
                << val_time << ")\n";
    } else {
      gai_host::out() << "train_time " << std::fixed << epoch_time << " s (fw " << fw_time << ", bw " << bw_time << ")\n";
    }
  }
  gai_host::out() << "Average training time per epoch: " << total_train_time / (double)num_epochs << " seconds. Throughput "
            << (double)num_epochs / total_train_time << " epoch/s\n";
}

// net.cpp:287-358. The frontier walk runs on the host (sequential, rand_r), the induced subgraph is built on the device; the feature rows
// of the kept vertices are gathered on the device from the resident feature matrix, the labels on the host.
template <typename L>
void Model<L>::subgraph_sampling(int) {
  if (num_subg_remain == 0) {
    for (int sid = 0; sid < num_subgraphs; sid++) {
      VertexSet sampled;
      // the reference seeds each walk with its OpenMP thread id (net.cpp:297-299); num_subgraphs == num_threads, static schedule: seed = sid
      sampler->select_vertices((index_t)subg_size, sampled, (unsigned)sid);
      if (subgs[sid]) { subgs[sid]->dealloc(); delete subgs[sid]; }
      subgs[sid] = new Graph(true);
      sampler->generateSubgraph(sampled, &subg_masks[(size_t)sid * num_samples], subgs[sid]);
    }
    num_subg_remain = num_subgraphs;
  }
  num_subg_remain--;
  const int sg_id = num_subg_remain;
  Graph* sg = subgs[sg_id];
  sg->degree_counting();
  subg_nv = (int)sg->size();
  // the subgraph is induced from the full graph, which already carries the self-loops of GCN / GAT: none are added (the reference's CPU
  // path adds none either; its GPU path would add a second one, net.cpp:325-327)
  if (!sg->device()) sg->copy_to_gpu();
  for (auto& y : layer_gconv) { y.update_dim_size((size_t)subg_nv); y.set_graph_ptr(sg); }
  if (use_l2norm) layer_l2norm->update_dim_size(subg_nv);
  if (use_dense) layer_dense->update_dim_size(subg_nv);
  layer_loss->update_dim_size(subg_nv);
  const mask_t* mk = &subg_masks[(size_t)sg_id * num_samples];
  std::vector<uint32_t> ids;
  ids.reserve(subg_nv);
  const size_t lw = is_sigmoid ? (size_t)num_cls : 1;
  labels_subg.resize((size_t)subg_nv * lw);
  for (int i = 0; i < num_samples; i++) {
    if (mk[i] != 1) continue;
    std::copy(labels.begin() + (size_t)i * lw, labels.begin() + (size_t)(i + 1) * lw, labels_subg.begin() + ids.size() * lw);
    ids.push_back((uint32_t)i);
  }
  assert((int)ids.size() == subg_nv);
  die_on(gai_memcpy_h2d(d_subg_ids, ids.data(), sizeof(uint32_t) * ids.size(), stream()), "gai_memcpy_h2d");
  die_on(gai_memcpy_h2d(d_labels_subg, labels_subg.data(), labels_subg.size(), stream()), "gai_memcpy_h2d");
  sync();  // the staging vectors go out of scope
  const int ld = (int)row_pitch(dim_init);
  die_on(gai_gather_rows((size_t)subg_nv, d_subg_ids, dim_init, d_input_features, ld, d_feats_subg, ld, stream()), "gai_gather_rows");
  layer_gconv[0].set_feat_in(d_feats_subg);
  layer_loss->set_labels_ptr(d_labels_subg);
}

template <typename L>
void Model<L>::use_full_graph() {
  for (auto& y : layer_gconv) { y.set_graph_ptr(full_graph); if (subg_size > 0) y.update_dim_size((size_t)num_samples); }
  if (subg_size > 0) {
    if (use_dense) layer_dense->update_dim_size(num_samples);
    if (use_l2norm) layer_l2norm->update_dim_size(num_samples);
    layer_loss->update_dim_size(num_samples);
    layer_gconv[0].set_feat_in(d_input_features);
    layer_loss->set_labels_ptr(d_labels);
  }
}

template class Model<GCN_layer>;
template class Model<GAT_layer>;
template class Model<SAGE_layer>;
