#include "gai_layers.h"
#include "gai_dist.h"
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <random>
#include <type_traits>

using gai_host::die_on;
using gai_host::stream;

// ---- helpers ------------------------------------------------------------------------------------------------------

void init_glorot(size_t dim_x, size_t dim_y, vec_t& weight, unsigned seed) {
  // Glorot & Bengio uniform(-r, r), r = sqrt(6/(fan_in+fan_out)); one draw per weight, row-major, from the C++ standard
  // library's default engine seeded with `seed` — the generator the reference uses on the host (math_functions.cpp:11-19),
  // so identical toolchains give identical initial weights. Never initialised on the device (cuRAND would differ).
  const float r = (float)std::sqrt(6.0 / (double)(dim_x + dim_y));
  std::default_random_engine engine(seed);
  std::uniform_real_distribution<float> uniform(-r, r);
  weight.resize(dim_x * dim_y);
  for (size_t i = 0; i < dim_x * dim_y; i++) weight[i] = uniform(engine);
}

float* float_malloc_device_zero(size_t n) {
  void* p = nullptr;
  die_on(gai_malloc(&p, sizeof(float) * (n ? n : 1)), "gai_malloc");
  die_on(gai_memset(p, 0, sizeof(float) * (n ? n : 1), stream()), "gai_memset");
  return reinterpret_cast<float*>(p);
}
void copy_float_to_device(size_t n, const float* src_h, float* dst_d) {
  die_on(gai_memcpy_h2d(dst_d, src_h, sizeof(float) * n, stream()), "gai_memcpy_h2d");
  die_on(gai_stream_sync(stream()), "gai_stream_sync");
}
void copy_float_to_host(size_t n, const float* src_d, float* dst_h) {
  die_on(gai_memcpy_d2h(dst_h, src_d, sizeof(float) * n, stream()), "gai_memcpy_d2h");
  die_on(gai_stream_sync(stream()), "gai_stream_sync");
}
static float* upload_glorot(size_t dx, size_t dy, unsigned seed) {
  vec_t w;
  init_glorot(dx, dy, w, seed);
  float* d = float_malloc_device_zero(dx * dy);
  copy_float_to_device(dx * dy, w.data(), d);
  return d;
}
static std::string shape(size_t x, size_t y, size_t z) { return std::to_string(x) + "x" + std::to_string(y) + "x" + std::to_string(z); }
// matmul(x,y,z,A,B,C,transA,transB,accum) of the reference (math_functions.cpp:142-171) with explicit row pitches
static void mm(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, bool ta = false,
               bool tb = false, bool accum = false, int flags = 0) {
  gai_host::OpScope sc("LINEAR", shape(x, y, z) + (ta ? " TA" : "") + (tb ? " TB" : ""),
                       4.0 * ((double)x * z + (double)z * y + (double)x * y * (accum ? 2 : 1)), 2.0 * (double)x * y * z);
  // tall outputs live in layer-owned buffers whose rows are padded to 4 floats: every epilogue store can be 128-bit
  const int padded = (!ta && ldc % 4 == 0 && ldc >= ceil4(y)) ? GAI_EPI_PADDED : 0;
  die_on(gai_matmul_ld(x, y, z, A, lda, B, ldb, C, ldc, ta, tb, accum, flags | padded, stream()), "gai_matmul_ld");
}
// C = A1·op(B1) + A2·op(B2) in one pass; mask != NULL folds the d_relu of the layer below into the epilogue
static void mm_kcat(size_t x, size_t y, size_t z1, const float* A1, size_t lda1, const float* B1, size_t z2, const float* A2, size_t lda2,
                    const float* B2, float* C, size_t ldc, bool tb, int flags, const float* mask, size_t ldmask, const uint32_t* mask_bits = nullptr,
                    uint32_t* relu_bits = nullptr) {
  const bool masked = mask || mask_bits;
  gai_host::OpScope sc("LINEAR", shape(x, y, z1) + "+" + std::to_string(z2) + (tb ? " TB" : "") + " kcat" + (mask_bits ? " bitmask" : (mask ? " mask" : "")),
                       4.0 * ((double)x * (z1 + z2) + (double)(z1 + z2) * y + (double)x * y * (mask ? 2 : 1)) + (mask_bits || relu_bits ? (double)x * y / 8 : 0),
                       2.0 * (double)x * y * (z1 + z2));
  const size_t ldb1 = tb ? z1 : y, ldb2 = tb ? z2 : y, ldw = bits_pitch(y);
  const int padded = (ldc % 4 == 0 && ldc >= ceil4(y) && (!mask || (ldmask % 4 == 0 && ldmask >= ceil4(y)))) ? GAI_EPI_PADDED : 0;
  const float* mptr = mask_bits ? reinterpret_cast<const float*>(mask_bits) : mask;
  die_on(gai_matmul_kcat(x, y, z1, A1, lda1, B1, ldb1, z2, A2, lda2, B2, ldb2, C, ldc, tb,
                         flags | padded | (masked ? GAI_EPI_MASK : 0) | (mask_bits ? GAI_EPI_BITMASK : 0), mptr, mask_bits ? ldw : ldmask, relu_bits, ldw,
                         stream()), "gai_matmul_kcat");
}
static void mm_mask(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, float* C, size_t ldc, bool tb, const float* mask,
                    size_t ldmask, const uint32_t* mask_bits = nullptr) {
  gai_host::OpScope sc("LINEAR", shape(x, y, z) + (tb ? " TB" : "") + (mask_bits ? " bitmask" : " mask"),
                       4.0 * ((double)x * z + (double)z * y + (mask_bits ? 1.0 : 2.0) * (double)x * y) + (mask_bits ? (double)x * y / 8 : 0), 2.0 * (double)x * y * z);
  const int padded = (ldc % 4 == 0 && ldc >= ceil4(y) && (mask_bits || (ldmask % 4 == 0 && ldmask >= ceil4(y)))) ? GAI_EPI_PADDED : 0;
  if (mask_bits)
    die_on(gai_matmul_mask(x, y, z, A, lda, B, tb ? z : y, C, ldc, tb, reinterpret_cast<const float*>(mask_bits), bits_pitch(y), padded | GAI_EPI_BITMASK,
                           stream()), "gai_matmul_mask");
  else
    die_on(gai_matmul_mask(x, y, z, A, lda, B, tb ? z : y, C, ldc, tb, mask, ldmask, padded, stream()), "gai_matmul_mask");
}
// C = ReLU(A·B) + the sign bits of C for the layer above
static void mm_relu_bits(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, float* C, size_t ldc, uint32_t* bits) {
  gai_host::OpScope sc("LINEAR", shape(x, y, z) + " relu+bits", 4.0 * ((double)x * z + (double)z * y + (double)x * y) + (double)x * y / 8, 2.0 * (double)x * y * z);
  const int padded = (ldc % 4 == 0 && ldc >= ceil4(y)) ? GAI_EPI_PADDED : 0;
  die_on(gai_matmul_relu_bits(x, y, z, A, lda, B, y, C, ldc, padded, bits, bits_pitch(y), stream()), "gai_matmul_relu_bits");
}
// C1 = A·B1, C2 = A·B2, A read once
static void mm_ncat(size_t x, size_t z, const float* A, size_t lda, size_t y, const float* B1, float* C1, size_t ldc1, const float* B2, float* C2,
                    size_t ldc2) {
  gai_host::OpScope sc("LINEAR", shape(x, y, z) + " ncat2", 4.0 * ((double)x * z + 2.0 * (double)z * y + 2.0 * (double)x * y), 4.0 * (double)x * y * z);
  const int padded = (ldc1 % 4 == 0 && ldc1 >= ceil4(y) && ldc2 % 4 == 0 && ldc2 >= ceil4(y)) ? GAI_EPI_PADDED : 0;
  die_on(gai_matmul_ncat(x, z, A, lda, y, B1, y, C1, ldc1, y, B2, y, C2, ldc2, padded, stream()), "gai_matmul_ncat");
}
// weight gradients of one layer in one pass over the shared operand: dW1 = A1^T·B, dW2 = A2^T·B (two_a) / dW1 = A^T·B1, dW2 = A^T·B2 (two_b)
static void wgrad_two_a(size_t n, size_t y, const float* B, size_t ldb, size_t x, const float* A1, size_t lda1, float* C1, const float* A2, size_t lda2,
                        float* C2) {
  gai_host::OpScope sc("LINEAR", shape(x, y, n) + " TA two_a", 4.0 * ((double)n * (2 * x + y) + 2.0 * (double)x * y), 4.0 * (double)x * y * n);
  die_on(gai_wgrad_two_a(n, y, B, ldb, x, A1, lda1, C1, y, x, A2, lda2, C2, y, stream()), "gai_wgrad_two_a");
}
static void wgrad_two_b(size_t n, size_t x, const float* A, size_t lda, size_t y, const float* B1, size_t ldb1, float* C1, const float* B2, size_t ldb2,
                        float* C2) {
  gai_host::OpScope sc("LINEAR", shape(x, y, n) + " TA two_b", 4.0 * ((double)n * (x + 2 * y) + 2.0 * (double)x * y), 4.0 * (double)x * y * n);
  die_on(gai_wgrad_two_b(n, x, A, lda, y, B1, ldb1, C1, y, y, B2, ldb2, C2, y, stream()), "gai_wgrad_two_b");
}
static void d_relu_rows(size_t rows, int F, float* grad, size_t ldg, const float* data, size_t ldd) {
  gai_host::OpScope sc("RELU", "d_relu n=" + std::to_string(rows * (size_t)F), 12.0 * rows * F, 0);
  // pitched rows carry zero padding on both sides, so the flat kernel over rows*pitch elements is the same op
  if (ldg == ldd) die_on(gai_d_relu(rows * ldg, grad, data, grad, stream()), "gai_d_relu");
  else die_on(gai_d_relu_ld(rows, F, grad, ldg, data, ldd, grad, ldg, stream()), "gai_d_relu_ld");
}
// algorithmic bytes of one aggregation call: gather model of SURVEY.md §8d
static double spmm_bytes(Graph& g, int F, int extra_per_edge = 0) {
  const double n = (double)g.size(), nnz = (double)g.sizeEdges();  // partitioned: this rank's master rows and their edges
  return 4.0 * (nnz * F + n * F + nnz * (1 + extra_per_edge) + (n + 1) + n);
}

// ---- adam ---------------------------------------------------------------------------------------------------------

void adam::update_gpu(const size_t n, const float* dW, float* W) {
  auto it = moments.find(W);
  if (it == moments.end()) it = moments.emplace(W, std::make_pair(float_malloc_device_zero(n), float_malloc_device_zero(n))).first;
  gai_host::OpScope sc("ADAM", "n=" + std::to_string(n), 28.0 * n, 0);
  die_on(gai_adam_update(n, dW, W, it->second.first, it->second.second, alpha, b1, b2, b1_t, b2_t, eps, stream()), "gai_adam_update");
  b1_t *= b1;
  b2_t *= b2;
}
void adam::reset() {
  for (auto& kv : moments) { gai_free(kv.second.first); gai_free(kv.second.second); }
  moments.clear();
}

// ---- aggregators --------------------------------------------------------------------------------------------------

void GCN_Aggregator::init(int len, int, int, float, float) { length = len; }
enum { SPMM_GCN = 0, SPMM_MEAN = 1, SPMM_MEAN_T = 2 };  // gai_spmm_rows_ex modes
// One aggregation call of any GCN / SAGE form. Partitioned graphs: the halo rows of `in` are fetched from their owners first — in one
// piece, or (wide matrices) in column blocks on the pull stream while the blocks already here are aggregated (Graph::halo_exchange_begin;
// columns are independent, so the result is bit-identical).
static void spmm_any(Graph& g, int mode, const char* name, int len, const float* in, size_t ld_in, float* out, size_t ld_out, int flags,
                     const float* addend, const uint32_t* mask_bits, const float* static_halo) {
  const uint32_t n = (uint32_t)g.size();
  const bool exchange = g.partitioned() && g.comm()->world() > 1 && !static_halo;
  const int nblk = exchange ? Graph::halo_block_count(len) : 1;
  const std::string tag = std::string(name) + " F=" + std::to_string(len) + (mask_bits ? " bitmask" : "");
  if (nblk <= 1) {
    const float* halo = static_halo ? static_halo : g.halo_exchange(in, len, ld_in);
    gai_host::OpScope sc("AGGR", tag, spmm_bytes(g, len) + (addend ? 4.0 * g.size() * len : 0) + (mask_bits ? g.size() * len / 8.0 : 0),
                         2.0 * g.sizeEdges() * len);
    die_on(gai_spmm_rows_ex(g.device(), mode, 0, n, len, nullptr, nullptr, in, (int)ld_in, out, (int)ld_out, flags, addend, mask_bits,
                            mask_bits ? (int)bits_pitch(len) : 0, halo, n, stream()), "gai_spmm_rows_ex");
    return;
  }
  const Graph::HaloBlocks hb = g.halo_exchange_begin(in, len, ld_in);
  for (int k = 0; k < hb.n; k++) {
    g.halo_wait_block(k);
    const int c0 = hb.col0[k], w = hb.ncol[k];
    gai_host::OpScope sc("AGGR", tag + " /" + std::to_string(hb.n),
                         spmm_bytes(g, w) + (addend ? 4.0 * g.size() * w : 0) + (mask_bits ? g.size() * w / 8.0 : 0), 2.0 * g.sizeEdges() * w);
    die_on(gai_spmm_rows_ex(g.device(), mode, 0, n, w, nullptr, nullptr, in + c0, (int)ld_in, out + c0, (int)ld_out,
                            flags | (k + 1 < hb.n ? GAI_SPMM_SHARE_SMS : 0) /* the next block's pull runs next to this kernel */, addend ? addend + c0 : nullptr,
                            mask_bits ? mask_bits + c0 / 32 : nullptr, mask_bits ? (int)bits_pitch(len) : 0, hb.halo ? hb.halo + c0 : nullptr, n, stream()),
           "gai_spmm_rows_ex(column block)");
  }
  g.halo_exchange_end();
}

void GCN_Aggregator::aggregate_ld(int len, Graph& g, const float* in, size_t ld_in, float* out, size_t ld_out, int flags, const float* addend,
                                  const float* static_halo) {
  spmm_any(g, SPMM_GCN, "gcn", len, in, ld_in, out, ld_out, flags, addend, nullptr, static_halo);
}
// the normalised adjacency is symmetric, so the derivative is the same product (gcn_aggregator.cpp:35-46)
void GCN_Aggregator::d_aggregate_ld(int len, Graph& g, const float* grad_in, size_t ld_in, float* grad_out, size_t ld_out, int flags, const float* addend,
                                    const uint32_t* mask_bits) {
  spmm_any(g, SPMM_GCN, "gcn", len, grad_in, ld_in, grad_out, ld_out, flags, addend, mask_bits, nullptr);
}
void GCN_Aggregator::aggregate(int len, Graph& g, const float* in, float* out) { aggregate_ld(len, g, in, len, out, len, GAI_EPI_NONE, nullptr); }
void GCN_Aggregator::d_aggregate(int len, Graph& g, const float*, const float* grad_in, float* grad_out) { aggregate(len, g, grad_in, grad_out); }

void SAGE_Aggregator::init(int len, int, int, float, float) { length = len; }
void SAGE_Aggregator::aggregate_ld(int len, Graph& g, const float* in, size_t ld_in, float* out, size_t ld_out, int flags, const float* addend,
                                   const float* static_halo) {
  spmm_any(g, SPMM_MEAN, "mean", len, in, ld_in, out, ld_out, flags, addend, nullptr, static_halo);
}
void SAGE_Aggregator::d_aggregate_ld(int len, Graph& g, const float* grad_in, size_t ld_in, float* grad_out, size_t ld_out, int flags, const float* addend,
                                     const uint32_t* mask_bits) {
  spmm_any(g, SPMM_MEAN_T, "meanT", len, grad_in, ld_in, grad_out, ld_out, flags, addend, mask_bits, nullptr);
}
void SAGE_Aggregator::aggregate(int len, Graph& g, const float* in, float* out) { aggregate_ld(len, g, in, len, out, len, GAI_EPI_NONE, nullptr); }
void SAGE_Aggregator::d_aggregate(int len, Graph& g, const float*, const float* grad_in, float* grad_out) {
  d_aggregate_ld(len, g, grad_in, len, grad_out, len, GAI_EPI_NONE, nullptr);
}

void GAT_Aggregator::init(int len, int, int ne, float lr, float drop_rate) {
  length = len;
  attn_drop = drop_rate;
  assert(attn_drop >= 0.f && attn_drop < 1.f);
  // the reference accepts the rate and ignores it on its CPU path (the attention-dropout lines of gat_aggregator.cpp:78-79,138-139 are
  // commented out): same here, with a note
  if (attn_drop > 0.f) std::cerr << "note: score_drop = " << attn_drop << " is accepted and ignored (as the reference CPU path does)\n";
  // Multi-head attention is an extension (the reference has one head, gat_layer.cpp:3-42; BASELINE.json configs[2] names 8): the row is
  // cut into `heads` blocks, the attention vectors keep their length (head h uses its block of them), scores become nnz x heads.
  if (const char* e = std::getenv("GAI_GAT_HEADS")) heads = std::atoi(e);
  if (heads < 1 || heads > 32 || (heads & (heads - 1)) != 0 || len % heads != 0 ||
      (heads > 1 && ((len / heads) % 4 != 0 || (((len / heads) / 4) & ((len / heads) / 4 - 1)) != 0 || len / heads > 128 || len > 512))) {
    std::cerr << "GAT_Aggregator: GAI_GAT_HEADS = " << heads << " does not divide a width of " << len
              << " into blocks the kernels take (power-of-two head count, 4 / 8 / 16 / ... / 128 columns per head, width <= 512)\n";
    std::exit(EXIT_FAILURE);
  }
  d_alpha_l = upload_glorot(len, 1, 2);  // seeds 2 / 3: gat_aggregator.cpp:11-12
  d_alpha_r = upload_glorot(len, 1, 3);
  d_alpha_lgrad = float_malloc_device_zero(len);
  d_alpha_rgrad = float_malloc_device_zero(len);
  d_temp_scores = float_malloc_device_zero((size_t)ne * heads);
  d_norm_scores = float_malloc_device_zero((size_t)ne * heads);
  d_scores_grad = float_malloc_device_zero((size_t)ne * heads);
  alpha_opt = new adam(lr);
}
void GAT_Aggregator::aggregate_fused(int len, Graph& g, const float* in, float* out, int flags, const float*) {
  gai_host::OpScope sc("ATTN_FWD", "gat F=" + std::to_string(len) + (heads > 1 ? " H=" + std::to_string(heads) : ""),
                       spmm_bytes(g, len, 2) + 4.0 * g.size() * len + 8.0 * g.sizeEdges() * (heads - 1), 2.0 * g.sizeEdges() * len);
  die_on(gai_gat_forward_heads_ld(g.device(), len, heads, in, row_pitch(len), d_alpha_l, d_alpha_r, epsilon, d_temp_scores, d_norm_scores, out,
                                  row_pitch(len), flags, stream()), "gai_gat_forward");
}
void GAT_Aggregator::aggregate(int len, Graph& g, const float* in, float* out) { aggregate_fused(len, g, in, out, GAI_EPI_NONE, nullptr); }
void GAT_Aggregator::d_aggregate(int len, Graph& g, const float* feat_in, const float* grad_in, float* grad_out) {
  gai_host::OpScope sc("ATTN_BWD", "gat F=" + std::to_string(len) + (heads > 1 ? " H=" + std::to_string(heads) : ""),
                       2.0 * spmm_bytes(g, len, 3) + 4.0 * g.size() * len + 12.0 * g.sizeEdges() * (heads - 1), 4.0 * g.sizeEdges() * len);
  die_on(gai_gat_backward_heads_ld(g.device(), len, heads, feat_in, row_pitch(len), grad_in, row_pitch(len), epsilon, d_temp_scores, d_norm_scores,
                                   d_scores_grad, d_alpha_lgrad, d_alpha_rgrad, grad_out, row_pitch(len), stream()), "gai_gat_backward");
}
void GAT_Aggregator::update_weights(optimizer*) {  // own optimiser, two calls (gat_aggregator.cpp:202-205)
  alpha_opt->update_gpu(length, d_alpha_lgrad, d_alpha_l);
  alpha_opt->update_gpu(length, d_alpha_rgrad, d_alpha_r);
}

// ---- graph_conv_layer ---------------------------------------------------------------------------------------------

template <typename A>
graph_conv_layer<A>::graph_conv_layer(int id, int nv, int din, int dout, Graph* g, bool act, bool concat, float lr, float feat_drop, float score_drop)
    : level_(id), num_samples(nv), dim_in(din), dim_out(dout), graph(g), is_act(act), is_bias(false), use_concat(concat),
      feat_dropout_rate(feat_drop), score_dropout_rate(score_drop) {
  assert(feat_dropout_rate >= 0.f && feat_dropout_rate < 1.f);
  assert(score_dropout_rate >= 0.f && score_dropout_rate < 1.f);
  feat_scale = 1.f / (1.f - feat_dropout_rate);
  const size_t n = (size_t)nv;
  d_W_neigh = upload_glorot(din, dout, 1);  // seeds: graph_conv_layer.cpp:13,18
  d_W_neigh_grad = float_malloc_device_zero((size_t)din * dout);
  if (concat) {
    d_W_self = upload_glorot(din, dout, 2);
    d_W_self_grad = float_malloc_device_zero((size_t)din * dout);
  }
  // row pitches: every per-vertex buffer, layer 0's input included (Model's device copy of the features), has rows padded to 4 floats
  ld_in = row_pitch(din);
  ld_out = row_pitch(dout);
  // temporaries: only what this layer's schedule touches (the reference allocates all of them unconditionally)
  const bool transform_first = din > dout || std::is_same<A, GAT_Aggregator>::value;  // GAT always transforms first
  transform_first_ = transform_first;
  // partitioned graph: a matrix an aggregation gathers from is registered with the peer group (same construction order on every rank):
  // the other ranks read their halo rows out of it (Graph::halo_exchange)
  const bool part = g->partitioned();
  if (part && ((size_t)nv != g->size() || std::is_same<A, GAT_Aggregator>::value || feat_drop > 0.f)) {
    std::cerr << "partitioned training supports GCN / SAGE layers without dropout over the rank's own masters\n";
    std::exit(EXIT_FAILURE);
  }
  auto gathered = [&](size_t pitch) { float* p = float_malloc_device_zero(n * pitch); g->register_gather_buffer(p); return p; };
  if (transform_first) d_out_temp = gathered(ld_out);                         // forward gathers X·W
  if (!transform_first) d_in_temp1 = float_malloc_device_zero(n * row_pitch(din));
  if (!transform_first && id > 0) d_in_temp = gathered(row_pitch(din));      // backward gathers G·W^T
  if (id > 0) feat_in = transform_first ? float_malloc_device_zero(n * ld_in) : gathered(ld_in);  // forward gathers the input
  grad_in = transform_first ? gathered(ld_out) : float_malloc_device_zero(n * ld_out);           // backward gathers the output gradient
  if (part) {  // per-rank partial weight gradients, summed over the ranks into the public ones at the end of backward()
    d_W_neigh_grad_local = float_malloc_device_zero((size_t)din * dout);
    g->comm()->register_buffer(d_W_neigh_grad_local);
    if (concat) { d_W_self_grad_local = float_malloc_device_zero((size_t)din * dout); g->comm()->register_buffer(d_W_self_grad_local); }
  } else {
    d_W_neigh_grad_local = d_W_neigh_grad;
    d_W_self_grad_local = d_W_self_grad;
  }
  if (feat_dropout_rate > 0.f) {  // dropout_mask + in_temp of the reference (graph_conv_layer.cpp:33-38)
    d_drop_in = float_malloc_device_zero(n * ld_in);
    void* mp = nullptr;
    die_on(gai_malloc(&mp, n * ld_in), "gai_malloc");
    d_dropout_mask = reinterpret_cast<uint8_t*>(mp);
  }
  // aggregate-first layers apply ReLU in a dense-transform epilogue, which also emits the sign bits the layer above masks with
  if (act && !transform_first) d_relu_bits = reinterpret_cast<uint32_t*>(float_malloc_device_zero(n * bits_pitch(dout)));
  optm = new adam(lr);
}

// in_data of the reference's forward (gcn_layer.cpp:15-19): the dropped-out input in the TRAIN phase, feat_in otherwise
template <typename A>
const float* graph_conv_layer<A>::forward_input() {
  if (!(feat_dropout_rate > 0.f && phase_ == net_phase::TRAIN)) return feat_in;
  gai_host::OpScope sc("DROPOUT", "fwd n=" + std::to_string((size_t)num_samples * ld_in), 9.0 * num_samples * ld_in, 0);
  die_on(gai_dropout((size_t)num_samples * ld_in, feat_dropout_rate, feat_scale, 0x5eedULL + (uint64_t)level_, dropout_calls++, feat_in, d_dropout_mask,
                     d_drop_in, stream()), "gai_dropout");
  return d_drop_in;
}
// grad_out *= mask * scale (gcn_layer.cpp:58-59)
template <typename A>
void graph_conv_layer<A>::backward_dropout(float* grad_out) {
  if (level_ == 0 || !(feat_dropout_rate > 0.f) || grad_out == nullptr) return;
  gai_host::OpScope sc("DROPOUT", "bwd n=" + std::to_string((size_t)num_samples * ld_in), 9.0 * num_samples * ld_in, 0);
  die_on(gai_d_dropout((size_t)num_samples * ld_in, feat_scale, grad_out, d_dropout_mask, grad_out, stream()), "gai_d_dropout");
}

template <typename A>
void graph_conv_layer<A>::reduce_weight_grads() {
  if (!graph->partitioned()) return;
  gai_host::OpScope sc("ALLREDUCE", "dW " + std::to_string(dim_in) + "x" + std::to_string(dim_out), 4.0 * dim_in * dim_out * graph->comm()->world(), 0);
  graph->comm()->all_reduce_sum(d_W_neigh_grad_local, (size_t)dim_in * dim_out, d_W_neigh_grad);
  if (use_concat) graph->comm()->all_reduce_sum(d_W_self_grad_local, (size_t)dim_in * dim_out, d_W_self_grad);
}

template <typename A>
bool graph_conv_layer<A>::can_mask_grad_out_bits() const {
  // aggregate-first GCN / SAGE layers end their backward with an aggregation, whose epilogue can apply a sign-bit mask
  return level_ > 0 && dim_in <= dim_out && !std::is_same<A, GAT_Aggregator>::value;
}

template <typename A>
bool graph_conv_layer<A>::can_mask_grad_out() const {
  // transform-first layers end their backward with grad_out = (...)·W^T, a dense transform whose epilogue can apply the mask;
  // aggregate-first layers end with an aggregation (plus, for SAGE, an accumulating transform)
  return level_ > 0 && (dim_in > dim_out || std::is_same<A, GAT_Aggregator>::value);
}

template <typename A>
float* graph_conv_layer<A>::weight_ptr(const std::string& name) {
  if (name == "W") return d_W_neigh;
  if (name == "W_grad") return d_W_neigh_grad;
  if (name == "W_self") return d_W_self;
  if (name == "W_self_grad") return d_W_self_grad;
  if (name == "feat_in") return feat_in;
  if (name == "grad_in") return grad_in;
  if (name == "out_temp") return d_out_temp;
  if (name == "in_temp1") return d_in_temp1;
  return nullptr;
}
template <typename A>
size_t graph_conv_layer<A>::weight_size(const std::string& name) {
  if (name == "W" || name == "W_grad") return (size_t)dim_in * dim_out;
  if (name == "W_self" || name == "W_self_grad") return use_concat ? (size_t)dim_in * dim_out : 0;
  if (name == "feat_in" || name == "in_temp1") return (size_t)num_samples * dim_in;
  if (name == "grad_in" || name == "out_temp") return (size_t)num_samples * dim_out;
  return 0;
}
template <typename A>
void graph_conv_layer<A>::tensor_layout(const std::string& name, size_t* cols, size_t* ld) {
  *cols = 0; *ld = 0;
  if (name == "feat_in") { *cols = dim_in; *ld = ld_in; }
  else if (name == "in_temp1") { *cols = dim_in; *ld = row_pitch(dim_in); }
  else if (name == "grad_in" || name == "out_temp") { *cols = dim_out; *ld = ld_out; }
}
template class graph_conv_layer<GCN_Aggregator>;
template class graph_conv_layer<SAGE_Aggregator>;
template class graph_conv_layer<GAT_Aggregator>;

// ---- GCN ----------------------------------------------------------------------------------------------------------

GCN_layer::GCN_layer(int id, int nv, int din, int dout, Graph* g, bool act, float lr, float fd, float sd)
    : graph_conv_layer(id, nv, din, dout, g, act, false, lr, fd, sd) {
  aggr.init(din < dout ? din : dout, nv);
}

void GCN_layer::forward(float* feat_out) {
  const size_t x = num_samples, y = dim_in, z = dim_out, ldt = row_pitch(y);
  const int relu = is_act ? GAI_EPI_RELU : GAI_EPI_NONE;
  const float* in_data = forward_input();
  if (y > z) {  // transform first: aggregate at the narrower width; ReLU rides the SpMM epilogue
    mm(x, z, y, in_data, ld_in, d_W_neigh, z, d_out_temp, ld_out);
    aggr.aggregate_ld((int)z, *graph, d_out_temp, ld_out, feat_out, ld_out, relu, nullptr);
  } else {      // aggregate first; ReLU rides the GEMM epilogue
    aggr.aggregate_ld((int)y, *graph, in_data, ld_in, d_in_temp1, ldt, GAI_EPI_NONE, nullptr, input_static_halo);
    if (is_act && d_relu_bits) mm_relu_bits(x, z, y, d_in_temp1, ldt, d_W_neigh, feat_out, ld_out, d_relu_bits);
    else mm(x, z, y, d_in_temp1, ldt, d_W_neigh, z, feat_out, ld_out, false, false, false, relu);
  }
}

void GCN_layer::backward(float* feat_out, float* grad_out) {
  const size_t x = num_samples, y = dim_in, z = dim_out, ldt = row_pitch(y);
  if (is_act && !grad_premasked) d_relu_rows(x, (int)z, grad_in, ld_out, feat_out, ld_out);
  if (y > z) {
    aggr.d_aggregate_ld((int)z, *graph, grad_in, ld_out, d_out_temp, ld_out, GAI_EPI_NONE, nullptr);
    if (level_ > 0) {
      if (mask_grad_out) mm_mask(x, y, z, d_out_temp, ld_out, d_W_neigh, grad_out, ld_in, true, feat_in, ld_in, mask_bits_in);
      else mm(x, y, z, d_out_temp, ld_out, d_W_neigh, z, grad_out, ld_in, false, true);
    }
    mm(y, z, x, feat_dropout_rate > 0.f ? d_drop_in : feat_in, ld_in, d_out_temp, ld_out, d_W_neigh_grad_local, z, true, false);  // gcn_layer.cpp:48-50
  } else {
    if (level_ > 0) {
      mm(x, y, z, grad_in, ld_out, d_W_neigh, z, d_in_temp, ldt, false, true);
      aggr.d_aggregate_ld((int)y, *graph, d_in_temp, ldt, grad_out, ld_in, GAI_EPI_NONE, nullptr, mask_grad_out ? mask_bits_in : nullptr);
    }
    mm(y, z, x, d_in_temp1, ldt, grad_in, ld_out, d_W_neigh_grad_local, z, true, false);
  }
  reduce_weight_grads();
  backward_dropout(grad_out);
}

void GCN_layer::update_weight(optimizer* opt) { opt->update_gpu((size_t)dim_in * dim_out, d_W_neigh_grad, d_W_neigh); }  // shared optimiser (gcn_layer.cpp:62-66)

// ---- SAGE ---------------------------------------------------------------------------------------------------------
// Same sums as sage_layer.cpp:5-53, regrouped so that every tall matrix is streamed once:
//   forward, aggregate first   out = ReLU([ÂX | X]·[W_n; W_s])                 one K-concatenated transform (reference: 2 sgemm, beta = 1)
//   forward, transform first   [T | S] = X·[W_n | W_s]; out = ReLU(ÂT + S)     one N-concatenated transform, self term added in the SpMM epilogue
//   backward, transform first  dT = Âᵀ·dOut;  [dW_s | dW_n] = Xᵀ·[dOut | dT];  dX = mask(dT·W_nᵀ + dOut·W_sᵀ)
//   backward, aggregate first  [dW_n; dW_s] = [ÂX | X]ᵀ·dOut;  dX = Âᵀ(dOut·W_nᵀ) + dOut·W_sᵀ  (self term first, neighbour term added by the SpMM)

SAGE_layer::SAGE_layer(int id, int nv, int din, int dout, Graph* g, bool act, float lr, float fd, float sd)
    : graph_conv_layer(id, nv, din, dout, g, act, true, lr, fd, sd) {
  aggr.init(din < dout ? din : dout, nv);
}

void SAGE_layer::forward(float* feat_out) {
  const size_t x = num_samples, y = dim_in, z = dim_out, ldt = row_pitch(y);
  const int relu = is_act ? GAI_EPI_RELU : GAI_EPI_NONE;
  const float* in_data = forward_input();
  if (y > z) {
    mm_ncat(x, y, in_data, ld_in, z, d_W_neigh, d_out_temp, ld_out, d_W_self, feat_out, ld_out);
    aggr.aggregate_ld((int)z, *graph, d_out_temp, ld_out, feat_out, ld_out, GAI_EPI_ADD | relu, feat_out);
  } else {
    aggr.aggregate_ld((int)y, *graph, in_data, ld_in, d_in_temp1, ldt, GAI_EPI_NONE, nullptr, input_static_halo);
    mm_kcat(x, z, y, d_in_temp1, ldt, d_W_neigh, y, in_data, ld_in, d_W_self, feat_out, ld_out, false, relu, nullptr, 0, nullptr,
            is_act ? d_relu_bits : nullptr);
  }
}

void SAGE_layer::backward(float* feat_out, float* grad_out) {
  const size_t x = num_samples, y = dim_in, z = dim_out, ldt = row_pitch(y);
  if (is_act && !grad_premasked) d_relu_rows(x, (int)z, grad_in, ld_out, feat_out, ld_out);
  const float* in_data = feat_dropout_rate > 0.f ? d_drop_in : feat_in;  // sage_layer.cpp:35-36
  if (y > z) {
    aggr.d_aggregate_ld((int)z, *graph, grad_in, ld_out, d_out_temp, ld_out, GAI_EPI_NONE, nullptr);
    wgrad_two_b(x, y, in_data, ld_in, z, grad_in, ld_out, d_W_self_grad_local, d_out_temp, ld_out, d_W_neigh_grad_local);
    if (level_ > 0)
      mm_kcat(x, y, z, d_out_temp, ld_out, d_W_neigh, z, grad_in, ld_out, d_W_self, grad_out, ld_in, true, 0,
              (mask_grad_out && !mask_bits_in) ? feat_in : nullptr, ld_in, mask_grad_out ? mask_bits_in : nullptr);
  } else {
    wgrad_two_a(x, z, grad_in, ld_out, y, d_in_temp1, ldt, d_W_neigh_grad_local, in_data, ld_in, d_W_self_grad_local);
    if (level_ > 0) {
      mm(x, y, z, grad_in, ld_out, d_W_neigh, z, d_in_temp, ldt, false, true);
      mm(x, y, z, grad_in, ld_out, d_W_self, z, grad_out, ld_in, false, true);
      aggr.d_aggregate_ld((int)y, *graph, d_in_temp, ldt, grad_out, ld_in, GAI_EPI_ADD, grad_out, mask_grad_out ? mask_bits_in : nullptr);
    }
  }
  reduce_weight_grads();
  backward_dropout(grad_out);
}

void SAGE_layer::update_weight(optimizer*) {  // the layer's own optimiser, neighbour then self (sage_layer.cpp:55-59)
  optm->update_gpu((size_t)dim_in * dim_out, d_W_neigh_grad, d_W_neigh);
  optm->update_gpu((size_t)dim_in * dim_out, d_W_self_grad, d_W_self);
}

// ---- GAT ----------------------------------------------------------------------------------------------------------

GAT_layer::GAT_layer(int id, int nv, int din, int dout, Graph* g, bool act, float lr, float fd, float sd)
    : graph_conv_layer(id, nv, din, dout, g, act, false, lr, fd, sd) {
  aggr.init(dout, nv, (int)g->sizeEdges(), lr, sd);
}

void GAT_layer::forward(float* feat_out) {
  const size_t x = num_samples, y = dim_in, z = dim_out;
  mm(x, z, y, forward_input(), ld_in, d_W_neigh, z, d_out_temp, ld_out);
  aggr.aggregate_fused((int)z, *graph, d_out_temp, feat_out, is_act ? GAI_EPI_RELU : GAI_EPI_NONE, nullptr);
}

void GAT_layer::backward(float* feat_out, float* grad_out) {
  const size_t x = num_samples, y = dim_in, z = dim_out;
  if (is_act && !grad_premasked) d_relu_rows(x, (int)z, grad_in, ld_out, feat_out, ld_out);
  aggr.d_aggregate((int)z, *graph, d_out_temp, grad_in, d_out_temp);  // dZ overwrites Z (gat_layer.cpp:33-36)
  if (level_ != 0) {
    if (mask_grad_out) mm_mask(x, y, z, d_out_temp, ld_out, d_W_neigh, grad_out, ld_in, true, feat_in, ld_in, mask_bits_in);
    else mm(x, y, z, d_out_temp, ld_out, d_W_neigh, z, grad_out, ld_in, false, true);
  }
  mm(y, z, x, feat_dropout_rate > 0.f ? d_drop_in : feat_in, ld_in, d_out_temp, ld_out, d_W_neigh_grad, z, true, false);  // gat_layer.cpp:32
  backward_dropout(grad_out);
}

void GAT_layer::update_weight(optimizer* opt) {
  opt->update_gpu((size_t)dim_in * dim_out, d_W_neigh_grad, d_W_neigh);
  aggr.update_weights(opt);
}

// ---- l2norm / dense / loss ------------------------------------------------------------------------------------------

l2norm_layer::l2norm_layer(int nv, int len) : num_samples(nv), dim(len) {
  feat_in = float_malloc_device_zero((size_t)nv * row_pitch(len));
  grad_in = float_malloc_device_zero((size_t)nv * row_pitch(len));
}
void l2norm_layer::forward(float* feat_out) {
  gai_host::OpScope sc("NORM", "l2norm", 8.0 * num_samples * dim, 0);
  die_on(gai_l2norm_ld(num_samples, dim, feat_in, row_pitch(dim), feat_out, row_pitch(dim), stream()), "gai_l2norm");
}
void l2norm_layer::backward(float* grad_out) {
  gai_host::OpScope sc("NORM", "d_l2norm", 12.0 * num_samples * dim, 0);
  die_on(gai_d_l2norm_ld(num_samples, dim, feat_in, row_pitch(dim), grad_in, row_pitch(dim), grad_out, row_pitch(dim), stream()), "gai_d_l2norm");
}

dense_layer::dense_layer(int nv, int in_len, int out_len, float lr) : dim_in(in_len), dim_out(out_len), num_samples(nv) {
  feat_in = float_malloc_device_zero((size_t)nv * row_pitch(in_len));
  grad_in = float_malloc_device_zero((size_t)nv * row_pitch(out_len));
  d_weight = upload_glorot(in_len, out_len, 1);  // dense_layer.cpp:33
  d_weight_grad = float_malloc_device_zero((size_t)in_len * out_len);
  optm = new adam(lr);
}
void dense_layer::forward(float* feat_out) { mm(num_samples, dim_out, dim_in, feat_in, row_pitch(dim_in), d_weight, dim_out, feat_out, row_pitch(dim_out)); }
void dense_layer::backward(float* grad_out) {
  mm(dim_in, dim_out, num_samples, feat_in, row_pitch(dim_in), grad_in, row_pitch(dim_out), d_weight_grad, dim_out, true, false);
  mm(num_samples, dim_in, dim_out, grad_in, row_pitch(dim_out), d_weight, dim_out, grad_out, row_pitch(dim_in), false, true);
  optm->update_gpu((size_t)dim_in * dim_out, d_weight_grad, d_weight);
}

loss_layer::loss_layer(int nv, int ncls, label_t* ptr) : num_samples(nv), num_cls(ncls), labels(ptr) {
  feat_in = float_malloc_device_zero((size_t)nv * row_pitch(ncls));
  feat_out = float_malloc_device_zero((size_t)nv * row_pitch(ncls));
  d_losses = float_malloc_device_zero(nv);
  d_stats = float_malloc_device_zero(4);
}

void loss_layer::set_partition(gai_host::Comm* comm) {
  comm_ = comm;
  comm->register_buffer(d_stats);
  d_stats_all = float_malloc_device_zero(4 * (size_t)comm->world());
}

void loss_layer::combine_stats(float* h) {
  if (!comm_) { copy_float_to_host(3, d_stats, h); return; }
  comm_->all_gather(d_stats, 4, d_stats_all);
  std::vector<float> all(4 * (size_t)comm_->world());
  copy_float_to_host(all.size(), d_stats_all, all.data());
  double loss = 0, correct = 0, count = 0;
  for (int q = 0; q < comm_->world(); q++) {
    const double c = all[4 * q + 2];
    if (c > 0) { loss += (double)all[4 * q] * c; correct += (double)all[4 * q + 1] * c; count += c; }
  }
  h[0] = count > 0 ? (float)(loss / count) : 0.f;
  h[1] = count > 0 ? (float)(correct / count) : 0.f;
  h[2] = (float)count;
}

void softmax_loss_layer::forward(size_t begin, size_t end, mask_t* masks) {
  if (comm_) die_on(gai_memset(d_stats, 0, 4 * sizeof(float), stream()), "gai_memset");  // a rank without rows of the range reports count 0
  // one pass over the logits also yields the loss mean and the accuracy get_prediction_loss() reports (forward_prop calls the two back to back)
  gai_host::OpScope sc("LOSS", "fwd+stats", 4.0 * (end - begin) * (2 * num_cls + 2), 0);
  die_on(gai_softmax_ce_forward_stats_ld(num_cls, begin, end, masks, labels, feat_in, row_pitch(num_cls), feat_out, row_pitch(num_cls), d_losses, d_stats,
                                         stream()), "gai_softmax_ce_forward_stats");
  stats_begin = begin; stats_end = end; stats_masks = masks; stats_valid = true;
}
void softmax_loss_layer::backward(size_t begin, size_t end, mask_t* masks, float* grad_out) {
  gai_host::OpScope sc("LOSS", "bwd", 8.0 * (end - begin) * num_cls, 0);
  if (begin == end) return;
  die_on(gai_softmax_ce_backward_ld(num_cls, begin, end, masks, labels, feat_out, row_pitch(num_cls), grad_out, row_pitch(num_cls),
                                    (uint64_t)(global_denom ? global_denom : end - begin), stream()), "gai_softmax_ce_backward");
}
acc_t softmax_loss_layer::get_prediction_loss(size_t begin, size_t end, size_t count, mask_t* masks) {
  if (!(stats_valid && stats_begin == begin && stats_end == end && stats_masks == masks)) {  // not preceded by forward() on the same rows
    gai_host::OpScope sc("LOSS", "reduce", 4.0 * (end - begin) * (num_cls + 1), 0);
    if (comm_) die_on(gai_memset(d_stats, 0, 4 * sizeof(float), stream()), "gai_memset");
    if (begin != end)
      die_on(gai_masked_loss_accuracy_ld(num_cls, begin, end, masks, labels, feat_in, row_pitch(num_cls), d_losses, d_stats, stream()),
             "gai_masked_loss_accuracy");
  }
  stats_valid = false;
  float h[3] = {0, 0, 0};
  combine_stats(h);
  (void)count;  // the reference asserts masked-row count == count; the count comes back as a float here
  last_acc = h[1];
  return h[0];
}
acc_t softmax_loss_layer::masked_accuracy(size_t begin, size_t end, mask_t* masks) {
  if (comm_) die_on(gai_memset(d_stats, 0, 4 * sizeof(float), stream()), "gai_memset");
  if (begin != end)
    die_on(gai_masked_loss_accuracy_ld(num_cls, begin, end, masks, labels, feat_in, row_pitch(num_cls), d_losses, d_stats, stream()),
           "gai_masked_loss_accuracy");
  stats_valid = false;
  float h[3] = {0, 0, 0};
  combine_stats(h);
  return h[1];
}

void sigmoid_loss_layer::forward(size_t begin, size_t end, mask_t* masks) {
  gai_host::OpScope sc("LOSS", "sigmoid fwd", 4.0 * (end - begin) * (2 * num_cls + 1) + (double)(end - begin) * num_cls, 0);
  die_on(gai_sigmoid_ce_forward_ld(num_cls, begin, end, masks, labels, feat_in, row_pitch(num_cls), feat_out, row_pitch(num_cls), d_losses, stream()),
         "gai_sigmoid_ce_forward");
}
void sigmoid_loss_layer::backward(size_t begin, size_t end, mask_t* masks, float* grad_out) {
  gai_host::OpScope sc("LOSS", "sigmoid bwd", 8.0 * (end - begin) * num_cls, 0);
  if (begin == end) return;
  die_on(gai_sigmoid_ce_backward_ld(num_cls, begin, end, masks, labels, feat_out, row_pitch(num_cls), grad_out, row_pitch(num_cls), (uint64_t)(end - begin),
                                    stream()), "gai_sigmoid_ce_backward");
}
acc_t sigmoid_loss_layer::get_prediction_loss(size_t begin, size_t end, size_t count, mask_t* masks) {
  {
    gai_host::OpScope sc("LOSS", "reduce", 4.0 * (end - begin), 0);
    die_on(gai_masked_loss_mean(begin, end, masks, d_losses, d_stats, stream()), "gai_masked_loss_mean");
  }
  float h[3] = {0, 0, 0};
  copy_float_to_host(3, d_stats, h);
  (void)count;
  return h[0];
}

float masked_accuracy_multi(int begin, int end, int, int num_classes, mask_t* masks, float* preds, label_t* ground_truth) {
  static thread_local float* d_f1 = nullptr;
  if (!d_f1) d_f1 = float_malloc_device_zero(4);
  die_on(gai_masked_f1_micro(num_classes, begin, end, masks, ground_truth, preds, row_pitch(num_classes), d_f1, stream()), "gai_masked_f1_micro");
  float h = 0.f;
  copy_float_to_host(1, d_f1, &h);
  return h;
}

// preds: the loss layer's feat_in (rows pitched to 4 floats, as every layer-owned buffer)
float masked_accuracy_single(int begin, int end, int, int num_classes, mask_t* masks, float* preds, label_t* ground_truth) {
  static thread_local float* scratch = nullptr;
  static thread_local size_t scratch_n = 0;
  if (scratch_n < (size_t)end + 4) {
    if (scratch) gai_free(scratch);
    scratch_n = (size_t)end + 4;
    scratch = float_malloc_device_zero(scratch_n);
  }
  die_on(gai_masked_loss_accuracy_ld(num_classes, begin, end, masks, ground_truth, preds, row_pitch(num_classes), scratch, scratch + end, stream()),
         "gai_masked_loss_accuracy");
  float h[3];
  copy_float_to_host(3, scratch + end, h);
  return h[1];
}
