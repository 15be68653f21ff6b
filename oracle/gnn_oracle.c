/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's GNN-layer hot path (oracle).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this. The product path
 * (graphaibench_b200/csrc + host) never calls it and has no CPU fallback.
 *
 * Each function restates one reference routine (chenxuhao/GraphAIBench @17834cb2, paths relative to the
 * reference root) in plain C with the SAME operation order and the same roundings, so that integer work
 * is bit-exact and the sparse fp32 work is bit-exact too (the reference does a rounded multiply followed
 * by a rounded add, accumulated sequentially in CSR edge order). Build: gcc -O2 -ffp-contract=off -fopenmp
 * (x86-64 SSE2 → IEEE binary32 arithmetic, no FMA contraction, no reassociation).
 *
 * Pinning: validated against the reference itself (oracle/_ref/libref_gnn.so, built from the reference's own
 * sources by oracle/build_ref.sh) in tests/test_oracle.py (test_restatement_against_live_reference), and against golden vectors generated from
 * that build and committed under tests/golden/ (tests/golden/make_golden.py). The one routine that is NOT
 * pinned bit-for-bit is orc_gemm: the reference calls cblas_sgemm from an unpinned third-party BLAS
 * (math_functions.cpp:148); orc_gemm accumulates in double and rounds once, which is at least as close to
 * the exact product as any fp32 BLAS — parity for dense transforms is to tolerance only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t index_t; /* include/gnn/global.h:75 */

/* ---- init_glorot: src/utilities/math_functions.cpp:11-19 -------------------------------------------
 * std::default_random_engine == minstd_rand0 (x <- 16807*x mod 2^31-1, x0 = seed) in libstdc++;
 * uniform_real_distribution<float>(a,b) == generate_canonical<float,24>() * (b-a) + a, and with
 * urng range 2^31-2 one draw suffices: canonical = float(x - 1) / float(2147483646.0L) (== 2^31). */
void orc_init_glorot(size_t dim_x, size_t dim_y, float* w, unsigned seed) {
  float init_range = (float)sqrt(6.0 / (double)(dim_x + dim_y));
  float a = -init_range, b = init_range;
  uint64_t x = seed % 2147483647u;
  if (x == 0) x = 1;
  const float denom = (float)2147483646.0L;
  for (size_t i = 0; i < dim_x * dim_y; i++) {
    x = (x * 16807ull) % 2147483647ull;
    float c = (float)(x - 1) / denom;
    if (c >= 1.0f) c = nextafterf(1.0f, 0.0f);
    w[i] = c * (b - a) + a;
  }
}

/* ---- LearningGraph::add_selfloop: include/gnn/lgraph.h:185-218 ------------------------------------
 * Inserts i into row i at its sorted position (rows are sorted, no pre-existing loops). colidx_out has
 * nnz+nv entries; rowptr_out[i] = rowptr[i] + i. */
void orc_add_selfloop(index_t nv, const index_t* rowptr, const index_t* colidx, index_t* rowptr_out, index_t* colidx_out) {
  for (index_t i = 0; i < nv; i++) {
    index_t start = rowptr[i], end = rowptr[i + 1];
    int inserted = 0;
    if (start == end) { colidx_out[start + i] = i; continue; }
    for (index_t e = start; e != end; e++) {
      index_t dst = colidx[e];
      if (!inserted) {
        if (i < dst) { inserted = 1; colidx_out[e + i] = i; colidx_out[e + i + 1] = dst; }
        else if (e + 1 == end) { inserted = 1; colidx_out[e + i + 1] = i; colidx_out[e + i] = dst; }
        else colidx_out[e + i] = dst;
      } else colidx_out[e + i + 1] = dst;
    }
  }
  for (index_t i = 0; i <= nv; i++) rowptr_out[i] = rowptr[i] + i;
}

/* ---- LearningGraph::compute_vertex_data: src/gnn/lgraph.cpp:22-34 ---------------------------------
 * temp = sqrtf(float(deg)); v = (float)(1.0 / (double)temp), 0 for isolated vertices. */
void orc_vertex_norm(index_t nv, const index_t* rowptr, float* vdata) {
  for (index_t v = 0; v < nv; v++) {
    index_t deg = rowptr[v + 1] - rowptr[v];
    float temp = sqrtf((float)deg);
    vdata[v] = (temp == 0.0f) ? 0.0f : (float)(1.0 / (double)temp);
  }
}

/* ---- LearningGraph::compute_edge_data: src/gnn/lgraph.cpp:6-20 ------------------------------------ */
void orc_edge_norm(index_t nv, const index_t* rowptr, const index_t* colidx, float* edata) {
  for (index_t i = 0; i < nv; i++) {
    float c_i = sqrtf((float)(rowptr[i + 1] - rowptr[i]));
    for (index_t e = rowptr[i]; e != rowptr[i + 1]; e++) {
      index_t j = colidx[e];
      float c_j = sqrtf((float)(rowptr[j + 1] - rowptr[j]));
      edata[e] = (c_i == 0.0f || c_j == 0.0f) ? 0.0f : (float)(1.0 / (double)(c_i * c_j));
    }
  }
}

/* ---- GCN_Aggregator::update_all: src/gnn/gconv/gcn_aggregator.cpp:48-77 ---------------------------
 * out_i = sum_e round(round(a_i*a_j) * in_j), accumulated in edge order; scale() then vadd_cpu()
 * (math_functions.cpp:266,336) are separate passes, hence two roundings and no FMA. */
void orc_spmm_gcn(index_t nv, const index_t* rowptr, const index_t* colidx, const float* vdata, int len, const float* in, float* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++) {
    float* o = out + (size_t)src * len;
    for (int k = 0; k < len; k++) o[k] = 0.0f;
    float a = vdata[src];
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++) {
      index_t dst = colidx[e];
      float b = a * vdata[dst];
      const float* x = in + (size_t)dst * len;
      for (int k = 0; k < len; k++) { float t = b * x[k]; o[k] = o[k] + t; }
    }
  }
}

/* ---- SAGE_Aggregator::aggregate / d_aggregate: src/gnn/gconv/sage_aggregator.cpp:7-54 --------------
 * forward b = (float)(1.0/float(deg_src)); transposed b = (float)(1.0/float(deg_dst)) per edge. */
void orc_spmm_mean(index_t nv, const index_t* rowptr, const index_t* colidx, int len, const float* in, float* out, int transposed) {
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++) {
    float* o = out + (size_t)src * len;
    for (int k = 0; k < len; k++) o[k] = 0.0f;
    float bs = (float)(1.0 / (double)(float)(rowptr[src + 1] - rowptr[src]));
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++) {
      index_t dst = colidx[e];
      float b = transposed ? (float)(1.0 / (double)(float)(rowptr[dst + 1] - rowptr[dst])) : bs;
      const float* x = in + (size_t)dst * len;
      for (int k = 0; k < len; k++) { float t = b * x[k]; o[k] = o[k] + t; }
    }
  }
}

/* ---- update_all with explicit per-edge scores: src/gnn/gconv/gat_aggregator.cpp:26-45
 *      (same loop as spmm(): math_functions.cpp:206-219) ------------------------------------------- */
void orc_spmm_edge(index_t nv, const index_t* rowptr, const index_t* colidx, const float* vals, int len, const float* in, float* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++) {
    float* o = out + (size_t)src * len;
    for (int k = 0; k < len; k++) o[k] = 0.0f;
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++) {
      const float* x = in + (size_t)colidx[e] * len;
      float s = vals[e];
      for (int k = 0; k < len; k++) { float t = s * x[k]; o[k] = o[k] + t; }
    }
  }
}

/* ---- symmetric_csr_transpose: src/utilities/math_functions.cpp:32-74 -------------------------------
 * For each edge e=(src,dst): idx = position of src in row dst (binary search); out[idx] = vals[e].
 * perm (optional) receives idx per edge: the e -> e^T permutation the GPU path builds once per graph. */
static int64_t orc_bsearch(const index_t* colidx, index_t key, int64_t begin, int64_t end) {
  int64_t l = begin, r = end - 1;
  while (r >= l) {
    int64_t mid = l + (r - l) / 2;
    index_t v = colidx[mid];
    if (v == key) return mid;
    if (v < key) l = mid + 1; else r = mid - 1;
  }
  return -1;
}
int orc_symmetric_transpose(index_t nv, const index_t* rowptr, const index_t* colidx, const float* vals, float* out, index_t* perm) {
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : bad)
  for (index_t src = 0; src < nv; src++) {
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++) {
      index_t dst = colidx[e];
      int64_t idx = orc_bsearch(colidx, src, rowptr[dst], rowptr[dst + 1]);
      if (idx < 0) { bad++; continue; }
      if (out) out[idx] = vals[e];
      if (perm) perm[e] = (index_t)idx;
    }
  }
  return bad;
}

/* ---- matmul: src/utilities/math_functions.cpp:142-171 (row-major cblas_sgemm, alpha=1, beta∈{0,1}) --
 * C[x×y] = op(A)·op(B) (+ C).  op(A) is x×z, op(B) is z×y.  Third-party BLAS, order unpinned: this
 * restatement accumulates in double and rounds once (see header). */
void orc_gemm(size_t x, size_t y, size_t z, const float* A, const float* B, float* C, int ta, int tb, int accum) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < x; i++) {
    double* acc = (double*)malloc(sizeof(double) * y);
    for (size_t j = 0; j < y; j++) acc[j] = 0.0;
    for (size_t k = 0; k < z; k++) {
      double a = ta ? A[k * x + i] : A[i * z + k];
      if (tb) { for (size_t j = 0; j < y; j++) acc[j] += a * (double)B[j * z + k]; }
      else { const float* b = B + k * y; for (size_t j = 0; j < y; j++) acc[j] += a * (double)b[j]; }
    }
    for (size_t j = 0; j < y; j++) C[i * y + j] = accum ? (float)((double)C[i * y + j] + acc[j]) : (float)acc[j];
    free(acc);
  }
}

/* ---- relu_cpu / d_relu_cpu: src/utilities/math_functions.cpp:442-463 ------------------------------ */
void orc_relu(size_t n, const float* in, float* out) {
  for (size_t i = 0; i < n; i++) out[i] = in[i] > 0.0f ? in[i] : 0.0f; /* std::max(in, 0): NaN -> NaN? max(a,b)=(a<b)?b:a -> in */
}
void orc_d_relu(size_t n, const float* in, const float* data, float* out) {
  for (size_t i = 0; i < n; i++) out[i] = data[i] > 0.0f ? in[i] : 0.0f;
}

/* ---- softmax: src/utilities/math_functions.cpp:485-494 -------------------------------------------- */
void orc_softmax(size_t n, const float* input, float* output) {
  float mx = input[0];
  for (size_t i = 1; i < n; i++) if (mx < input[i]) mx = input[i];
  float denom = 0.0f;
  for (size_t i = 0; i < n; i++) { output[i] = expf(input[i] - mx); denom += output[i]; }
  for (size_t i = 0; i < n; i++) output[i] /= denom;
}

/* ---- d_softmax, non-AVX512 branch (the default build): src/utilities/math_functions.cpp:496-514 ----
 * dy[i] = dot(dp, df_i), df_i[j] = (j==i) ? p[i]*(1-p[i]) : -p[j]*p[i]; O(n^2), sequential dot. */
void orc_d_softmax(int n, const float* p, const float* dp, float* dy) {
  for (int i = 0; i < n; i++) {
    float sum = 0.0f;
    for (int j = 0; j < n; j++) {
      float df = (j == i) ? p[i] * (1.0f - p[i]) : -p[j] * p[i];
      sum += dp[j] * df;
    }
    dy[i] = sum;
  }
}
/* AVX512 branch of the same routine (math_functions.cpp:497-504), O(n): used for hub rows in large cases. */
void orc_d_softmax_closed(int n, const float* p, const float* dp, float* dy) {
  float s = 0.0f;
  for (int i = 0; i < n; i++) s += p[i] * dp[i];
  for (int i = 0; i < n; i++) {
    float x = (float)((double)p[i] * (1.0 - (double)p[i]) * (double)dp[i]);
    dy[i] = x - (s - p[i] * dp[i]) * p[i];
  }
}

/* ---- softmax_loss_layer::forward/backward/get_prediction_loss: src/layers/softmax_loss_layer.cpp:4-55;
 *      cross_entropy: math_functions.cpp:531-542; masked_accuracy_single + argmax: :79-92,:129-139 ----
 * probs/losses written for masked rows in [begin,end); grad = (p - onehot)/(end-begin) (range length, not
 * count) on the same rows, other rows untouched. Returns mean loss over masked rows; *acc = accuracy. */
float orc_softmax_loss(int ncls, const float* logits, const uint8_t* labels, const uint8_t* masks, size_t begin, size_t end,
                       float* probs, float* losses, float* grad, float* acc) {
  float total = 0.0f, correct = 0.0f;
  size_t count = 0;
  for (size_t i = begin; i < end; i++) {
    if (masks && masks[i] != 1) continue;
    const float* x = logits + (size_t)ncls * i;
    float* p = probs + (size_t)ncls * i;
    orc_softmax(ncls, x, p);
    float pl = p[labels[i]];
    float loss = 0.0f;
    loss -= 1.0f * ((pl == 0.0f) ? logf(1e-10f) : logf(pl));
    losses[i] = loss;
    total += loss;
    count++;
    if (grad) {
      for (int j = 0; j < ncls; j++)
        grad[(size_t)ncls * i + j] = (float)(((double)p[j] - (labels[i] == j ? 1.0 : 0.0)) / (double)(end - begin));
    }
    int am = -1; float mx = -INFINITY;
    for (int j = 0; j < ncls; j++) if (x[j] > mx) { mx = x[j]; am = j; }
    if (am == (int)labels[i]) correct += 1.0f;
  }
  if (acc) *acc = correct / (float)count;
  return count ? total / (float)count : 0.0f;
}

/* ---- sigmoid_loss_layer::forward/backward/get_prediction_loss (src/layers/sigmoid_loss_layer.cpp:4-55), sigmoid and
 * sigmoid_cross_entropy (src/utilities/math_functions.cpp:517-521,553-559, with their mixed float/double arithmetic) and
 * masked_accuracy_multi = micro-F1 over (row, class) pairs at threshold 0.5 (math_functions.cpp:94-97,580-623).
 * labels: [nv x ncls] multi-hot. grad = (p - y) / (float)(end - begin) on masked rows, others untouched. Returns the mean loss. */
float orc_sigmoid_loss(int ncls, const float* logits, const uint8_t* labels, const uint8_t* masks, size_t begin, size_t end,
                       float* probs, float* losses, float* grad, float* f1) {
  float total = 0.0f;
  size_t count = 0;
  long tp = 0, fp = 0, fn = 0;
  for (size_t i = begin; i < end; i++) {
    if (masks && masks[i] != 1) continue;
    const float* x = logits + (size_t)ncls * i;
    const uint8_t* y = labels + (size_t)ncls * i;
    float* p = probs + (size_t)ncls * i;
    float loss = 0.0f;
    for (int j = 0; j < ncls; j++) {
      p[j] = (float)(1. / (1. + expf(-x[j])));
      loss -= x[j] * ((float)y[j] - (x[j] >= 0.)) - logf((float)(1. + expf((float)(x[j] - 2. * x[j] * (x[j] >= 0.)))));
    }
    losses[i] = loss;
    total += loss;
    count++;
    if (grad)
      for (int j = 0; j < ncls; j++) grad[(size_t)ncls * i + j] = (p[j] - (float)y[j]) / (float)(end - begin);
    for (int j = 0; j < ncls; j++) {
      if (y[j] == 1 && p[j] > 0.5) tp++;
      else if (y[j] == 0 && p[j] > 0.5) fp++;
      else if (y[j] == 1 && p[j] <= 0.5) fn++;
    }
  }
  if (f1) {
    const double prec = tp + fp > 0 ? (double)tp / (double)(tp + fp) : 0., rec = tp + fn > 0 ? (double)tp / (double)(tp + fn) : 0.;
    *f1 = (float)(rec + prec > 0. ? 2. * (rec * prec) / (rec + prec) : 0.);
  }
  return count ? total / (float)count : 0.0f;
}

/* ---- adam::update: src/utilities/optimizer.cpp:22-35 (eps inside the sqrt; b1_t/b2_t advance per call) */
void orc_adam(size_t n, const float* dW, float* W, float* m, float* v, float alpha, float b1, float b2, float* b1_t, float* b2_t, float eps) {
  for (size_t i = 0; i < n; i++) {
    m[i] = b1 * m[i] + (1.0f - b1) * dW[i];
    v[i] = b2 * v[i] + (1.0f - b2) * dW[i] * dW[i];
    W[i] -= alpha * (m[i] / (1.0f - *b1_t)) / sqrtf((v[i] / (1.0f - *b2_t)) + eps);
  }
  *b1_t *= b1;
  *b2_t *= b2;
}

/* ---- l2norm_layer::forward/backward: src/layers/l2norm_layer.cpp:19-64 ---------------------------- */
void orc_l2norm(int n, int dim, const float* in, float* out) {
  for (int i = 0; i < n; i++) {
    const float* x = in + (size_t)i * dim;
    float sum = 0.0f;
    for (int j = 0; j < dim; j++) sum += x[j] * x[j];
    sum = (sum < 1.0e-12) ? (float)1.0e-12 : sum;
    sum = sqrtf(sum);
    for (int j = 0; j < dim; j++) out[(size_t)i * dim + j] = x[j] / sum;
  }
}
void orc_d_l2norm(int n, int dim, const float* feat_in, const float* grad_in, float* grad_out) {
  for (int i = 0; i < n; i++) {
    const float* x = feat_in + (size_t)i * dim;
    const float* g = grad_in + (size_t)i * dim;
    float coef0 = 0.0f, sum_x2 = 0.0f;
    for (int j = 0; j < dim; j++) { sum_x2 += powf(x[j], 2.0f); coef0 -= x[j] * g[j]; }
    sum_x2 = (sum_x2 < 1.0e-12) ? (float)1.0e-12 : sum_x2;
    float coef1 = powf(sum_x2, -1.5f);
    for (int j = 0; j < dim; j++) grad_out[(size_t)i * dim + j] = x[j] * coef0 * coef1 + g[j] * sum_x2 * coef1;
  }
}

/* ---- GAT_Aggregator::aggregate: src/gnn/gconv/gat_aggregator.cpp:57-97 ----------------------------
 * temp_scores[e] = dot(alpha_l, z_src) + dot(alpha_r, z_dst); scores = LeakyReLU_0.2; row softmax →
 * norm_scores; out = update_all(norm_scores, z). dot(): math_functions.cpp:100-103 (sequential). */
static float orc_dot(int n, const float* x, const float* y) {
  float s = 0.0f;
  for (int i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}
void orc_gat_forward(index_t nv, const index_t* rowptr, const index_t* colidx, int len, const float* alpha_l, const float* alpha_r,
                     float slope, const float* z, float* temp_scores, float* scores, float* norm_scores, float* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++) {
    index_t b = rowptr[src], e_end = rowptr[src + 1];
    float ss = orc_dot(len, alpha_l, z + (size_t)src * len);
    for (index_t e = b; e != e_end; e++) {
      float ds = orc_dot(len, alpha_r, z + (size_t)colidx[e] * len);
      temp_scores[e] = ss + ds;
      scores[e] = temp_scores[e] > 0.0f ? temp_scores[e] : slope * temp_scores[e];
    }
    if (e_end > b) orc_softmax(e_end - b, scores + b, norm_scores + b);
  }
  orc_spmm_edge(nv, rowptr, colidx, norm_scores, len, z, out);
}

/* ---- GAT_Aggregator::d_aggregate: src/gnn/gconv/gat_aggregator.cpp:99-200 -------------------------
 * (1) SDDMM dS[e] = dot(g_src, z_dst); (2) per row softmax-bwd into scores[], LeakyReLU-bwd, alpha grads
 * accumulated over rows in ascending order (the reference's single-thread order; with several OpenMP
 * threads the reference's own summation order varies run to run); (3) transpose norm_scores on the symmetric
 * pattern; (4) dZ = update_all(scores^T, G). closed_form selects the AVX512 d_softmax branch. */
void orc_gat_backward(index_t nv, const index_t* rowptr, const index_t* colidx, int len, float slope, const float* z, const float* grad_in,
                      const float* temp_scores, const float* norm_scores, float* scores, float* norm_scores_grad,
                      float* alpha_lgrad, float* alpha_rgrad, float* grad_out, int closed_form) {
  index_t nnz = rowptr[nv];
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++)
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++)
      norm_scores_grad[e] = orc_dot(len, grad_in + (size_t)src * len, z + (size_t)colidx[e] * len);
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++) {
    index_t b = rowptr[src], n = rowptr[src + 1] - b;
    if (n == 0) continue;
    if (closed_form) orc_d_softmax_closed((int)n, norm_scores + b, norm_scores_grad + b, scores + b);
    else orc_d_softmax((int)n, norm_scores + b, norm_scores_grad + b, scores + b);
  }
  float* sum_l = (float*)calloc(len, sizeof(float));
  float* sum_r = (float*)calloc(len, sizeof(float));
  for (index_t src = 0; src < nv; src++) {
    float ssg = 0.0f;
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++) {
      float tsg = scores[e] * (temp_scores[e] > 0.0f ? 1.0f : slope);
      const float* x = z + (size_t)colidx[e] * len;
      for (int k = 0; k < len; k++) { float t = tsg * x[k]; sum_r[k] = t + sum_r[k]; }
      ssg += tsg;
    }
    const float* x = z + (size_t)src * len;
    for (int k = 0; k < len; k++) { float t = ssg * x[k]; sum_l[k] = t + sum_l[k]; }
  }
  for (int k = 0; k < len; k++) { alpha_lgrad[k] = 0.0f + sum_l[k]; alpha_rgrad[k] = 0.0f + sum_r[k]; }
  free(sum_l); free(sum_r);
  float* st = (float*)malloc(sizeof(float) * (nnz ? nnz : 1));
  orc_symmetric_transpose(nv, rowptr, colidx, norm_scores, st, NULL);
  orc_spmm_edge(nv, rowptr, colidx, st, len, grad_in, grad_out);
  free(st);
}

/* ---- Multi-head attention: an EXTENSION the reference does not have (single head everywhere, gat_layer.cpp:3-42); BASELINE.json
 * configs[2] names 8 heads. Defined so that heads == 1 is the reference path above, operation for operation (tests pin that bit for
 * bit): the feature row is cut into `heads` blocks of D = len / heads columns, head h uses alpha_l / alpha_r entries [hD, (h+1)D)
 * (same parameter count), has its own scores / row softmax, and aggregates its own block of columns (heads concatenated).
 * Score arrays are edge-major: x[e * heads + h]. */
void orc_gat_forward_heads(index_t nv, const index_t* rowptr, const index_t* colidx, int len, int heads, const float* alpha_l,
                           const float* alpha_r, float slope, const float* z, float* temp_scores, float* scores, float* norm_scores, float* out) {
  const int D = len / heads;
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++) {
    index_t b = rowptr[src], e_end = rowptr[src + 1];
    size_t n = e_end - b;
    float* col = (float*)malloc(sizeof(float) * 2 * (n ? n : 1));
    for (int h = 0; h < heads; h++) {
      float ss = orc_dot(D, alpha_l + h * D, z + (size_t)src * len + h * D);
      for (index_t e = b; e != e_end; e++) {
        float ds = orc_dot(D, alpha_r + h * D, z + (size_t)colidx[e] * len + h * D);
        float t = ss + ds;
        temp_scores[(size_t)e * heads + h] = t;
        col[e - b] = scores[(size_t)e * heads + h] = t > 0.0f ? t : slope * t;
      }
      if (n) {
        orc_softmax(n, col, col + n);
        for (index_t e = b; e != e_end; e++) norm_scores[(size_t)e * heads + h] = col[n + (e - b)];
      }
    }
    free(col);
    float* o = out + (size_t)src * len;
    for (int k = 0; k < len; k++) o[k] = 0.0f;
    for (index_t e = b; e != e_end; e++) {
      const float* x = z + (size_t)colidx[e] * len;
      for (int k = 0; k < len; k++) { float t = norm_scores[(size_t)e * heads + k / D] * x[k]; o[k] = o[k] + t; }
    }
  }
}

void orc_gat_backward_heads(index_t nv, const index_t* rowptr, const index_t* colidx, int len, int heads, float slope, const float* z,
                            const float* grad_in, const float* temp_scores, const float* norm_scores, float* scores, float* norm_scores_grad,
                            float* alpha_lgrad, float* alpha_rgrad, float* grad_out, int closed_form) {
  const int D = len / heads;
  index_t nnz = rowptr[nv];
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++)
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++)
      for (int h = 0; h < heads; h++)
        norm_scores_grad[(size_t)e * heads + h] = orc_dot(D, grad_in + (size_t)src * len + h * D, z + (size_t)colidx[e] * len + h * D);
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++) {
    index_t b = rowptr[src], n = rowptr[src + 1] - b;
    if (n == 0) continue;
    float* col = (float*)malloc(sizeof(float) * 3 * n);
    for (int h = 0; h < heads; h++) {
      for (index_t i = 0; i < n; i++) { col[i] = norm_scores[(size_t)(b + i) * heads + h]; col[n + i] = norm_scores_grad[(size_t)(b + i) * heads + h]; }
      if (closed_form) orc_d_softmax_closed((int)n, col, col + n, col + 2 * n);
      else orc_d_softmax((int)n, col, col + n, col + 2 * n);
      for (index_t i = 0; i < n; i++) scores[(size_t)(b + i) * heads + h] = col[2 * n + i];
    }
    free(col);
  }
  float* sum_l = (float*)calloc(len, sizeof(float));
  float* sum_r = (float*)calloc(len, sizeof(float));
  float* ssg = (float*)malloc(sizeof(float) * heads);
  for (index_t src = 0; src < nv; src++) {
    for (int h = 0; h < heads; h++) ssg[h] = 0.0f;
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++) {
      const float* x = z + (size_t)colidx[e] * len;
      for (int h = 0; h < heads; h++) {
        float tsg = scores[(size_t)e * heads + h] * (temp_scores[(size_t)e * heads + h] > 0.0f ? 1.0f : slope);
        for (int k = h * D; k < (h + 1) * D; k++) { float t = tsg * x[k]; sum_r[k] = t + sum_r[k]; }
        ssg[h] += tsg;
      }
    }
    const float* x = z + (size_t)src * len;
    for (int k = 0; k < len; k++) { float t = ssg[k / D] * x[k]; sum_l[k] = t + sum_l[k]; }
  }
  for (int k = 0; k < len; k++) { alpha_lgrad[k] = 0.0f + sum_l[k]; alpha_rgrad[k] = 0.0f + sum_r[k]; }
  free(sum_l); free(sum_r); free(ssg);
  /* dZ = P^T·G per head: transpose each head's scores on the symmetric pattern */
  index_t* perm = (index_t*)malloc(sizeof(index_t) * (nnz ? nnz : 1));
  orc_symmetric_transpose(nv, rowptr, colidx, NULL, NULL, perm);   /* perm[e] = position of the reverse edge (an involution) */
#pragma omp parallel for schedule(dynamic, 64)
  for (index_t src = 0; src < nv; src++) {
    float* o = grad_out + (size_t)src * len;
    for (int k = 0; k < len; k++) o[k] = 0.0f;
    for (index_t e = rowptr[src]; e != rowptr[src + 1]; e++) {
      const float* x = grad_in + (size_t)colidx[e] * len;
      const float* w = norm_scores + (size_t)perm[e] * heads;
      for (int k = 0; k < len; k++) { float t = w[k / D] * x[k]; o[k] = o[k] + t; }
    }
  }
  free(perm);
}

/* ---- PartitionedGraph::edgecut_induced_partition1D + generate_induced_subgraph:
 *      src/partitioner/graph_partition.cc:70-178 (rowptr is int64 there: include/common.h eidType) -----
 * Partition `part` of `nparts`: masters [S*part, min(S*(part+1), nv)), S = ceil(nv/nparts); vertex set =
 * masters ∪ their neighbours, local ids in ascending global id; induced CSR over the set.
 * Call once with idx_map==NULL to get sizes (returns m = |set|, *ne_out), then with buffers. */
int64_t orc_partition1d(index_t nv, const int64_t* rowptr, const index_t* colidx, int nparts, int part, index_t* idx_map,
                        int64_t* sub_rowptr, index_t* sub_colidx, int64_t* ne_out, index_t* local_begin, index_t* local_end) {
  int64_t S = nv / nparts; if (nv % nparts != 0) S++;
  int64_t bv = S * part, ev = bv + S; if (ev > nv) ev = nv;
  uint8_t* mask = (uint8_t*)calloc(nv, 1);
  for (int64_t v = bv; v < ev; v++) { mask[v] = 1; for (int64_t e = rowptr[v]; e < rowptr[v + 1]; e++) mask[colidx[e]] = 1; }
  index_t* new_id = (index_t*)malloc(sizeof(index_t) * nv);
  int64_t m = 0;
  for (index_t v = 0; v < nv; v++) if (mask[v]) { new_id[v] = (index_t)m; if (idx_map) idx_map[m] = v; m++; }
  int64_t ne = 0;
  for (index_t v = 0; v < nv; v++) if (mask[v]) {
    if (sub_rowptr) sub_rowptr[new_id[v]] = ne;
    for (int64_t e = rowptr[v]; e < rowptr[v + 1]; e++) if (mask[colidx[e]]) { if (sub_colidx) sub_colidx[ne] = new_id[colidx[e]]; ne++; }
  }
  if (sub_rowptr) sub_rowptr[m] = ne;
  if (ne_out) *ne_out = ne;
  if (local_begin && bv < ev) { *local_begin = new_id[bv]; *local_end = new_id[ev - 1] + 1; }
  free(mask); free(new_id);
  return m;
}
