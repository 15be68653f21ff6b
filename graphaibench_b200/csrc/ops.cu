// Elementwise ops, row-wise L2 normalisation, masked softmax cross-entropy, Adam and the matmul entry point.
// Replaces relu_gpu/d_relu_gpu/init_const_gpu/l2norm/d_l2norm/softmax_cross_entropy/d_softmax_cross_entropy/
// masked_avg_loss/masked_accuracy_single (src/utilities/math_functions.cu) and adam::update_gpu (optimizer.cu:5-36).
#include "gai_internal.cuh"

namespace gai {
int gemm_simt(size_t M, size_t N, size_t K, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int ta, int tb,
              int accum, int flags, cudaStream_t st);
int gemm_tc(size_t M, size_t N, size_t K, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int ta, int tb,
            int accum, int flags, int passes, cudaStream_t st);  // returns GAI_ERR_UNSUPPORTED when the shape is not taken
int gemm_tc_wgrad(size_t Kx, size_t My, size_t nrows, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int accum,
                  int flags, int passes, cudaStream_t st);
int gemm_tc_cat(const GemmCat& q, int passes, cudaStream_t st);
int gemm_tc_wgrad_cat(const WgradCat& q, int passes, cudaStream_t st);
static int g_gemm_mode = 0;
}  // namespace gai

namespace {

inline unsigned grid_for(size_t n, int threads, int per_thread = 1) {
  size_t blocks = (n + (size_t)threads * per_thread - 1) / ((size_t)threads * per_thread);
  size_t cap = (size_t)gai::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks ? blocks : 1);
}

__global__ void relu_kernel(size_t n, const float* __restrict__ in, float* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = n / 4;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0) {
    for (size_t k = i; k < n4; k += stride) {
      float4 v = reinterpret_cast<const float4*>(in)[k];
      v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f; v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
      reinterpret_cast<float4*>(out)[k] = v;
    }
    for (size_t k = n4 * 4 + i; k < n; k += stride) out[k] = in[k] > 0.f ? in[k] : 0.f;
  } else {
    for (size_t k = i; k < n; k += stride) out[k] = in[k] > 0.f ? in[k] : 0.f;
  }
}

// out = data > 0 ? grad : 0   (d_relu_cpu, math_functions.cpp:453-463)
__global__ void d_relu_kernel(size_t n, const float* __restrict__ grad, const float* __restrict__ data, float* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = n / 4;
  if ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(data) | reinterpret_cast<uintptr_t>(out)) % 16 == 0) {
    for (size_t k = i; k < n4; k += stride) {
      const float4 g = reinterpret_cast<const float4*>(grad)[k];
      const float4 d = reinterpret_cast<const float4*>(data)[k];
      float4 v;
      v.x = d.x > 0.f ? g.x : 0.f; v.y = d.y > 0.f ? g.y : 0.f; v.z = d.z > 0.f ? g.z : 0.f; v.w = d.w > 0.f ? g.w : 0.f;
      reinterpret_cast<float4*>(out)[k] = v;
    }
    for (size_t k = n4 * 4 + i; k < n; k += stride) out[k] = data[k] > 0.f ? grad[k] : 0.f;
  } else {
    for (size_t k = i; k < n; k += stride) out[k] = data[k] > 0.f ? grad[k] : 0.f;
  }
}

__global__ void d_relu_ld_kernel(size_t rows, int F, const float* __restrict__ grad, size_t ldg, const float* __restrict__ data, size_t ldd,
                                 float* __restrict__ out, size_t ldo) {
  const size_t total = rows * (size_t)F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / F, c = i % F;
    out[r * ldo + c] = data[r * ldd + c] > 0.f ? grad[r * ldg + c] : 0.f;
  }
}

// sign bits of a row-major matrix: bits[r, c / 32] bit (c % 32) = data[r, c] > 0 (one warp per word group; fallback for shapes the
// tensor-core epilogue does not take), and the matching mask: out = bit ? grad : 0
__global__ void sign_bits_kernel(size_t rows, int F, const float* __restrict__ data, size_t ldd, uint32_t* __restrict__ bits, size_t ldb) {
  const int nw = (F + 31) / 32;
  const size_t total = rows * (size_t)nw;
  const int lane = threadIdx.x & 31;
  for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += ((size_t)gridDim.x * blockDim.x) >> 5) {
    const size_t r = w / nw;
    const int c = (int)(w % nw) * 32 + lane;
    const unsigned b = __ballot_sync(0xffffffffu, c < F && data[r * ldd + c] > 0.f);
    if (lane == 0) bits[r * ldb + w % nw] = b;
  }
}
__global__ void d_relu_bits_kernel(size_t rows, int F, float* __restrict__ grad, size_t ldg, const uint32_t* __restrict__ bits, size_t ldb) {
  const size_t total = rows * (size_t)F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / F, c = i % F;
    if (!((bits[r * ldb + (c >> 5)] >> (c & 31)) & 1u)) grad[r * ldg + c] = 0.f;
  }
}

// dropout / d_dropout (math_functions.cpp:404-440): out = in * (float)mask * scale, mask ~ Bernoulli(1 - rate) per element, redrawn on
// every call. The reference draws from a /dev/urandom-seeded boost mt19937 (random.cpp:6-20), i.e. it is not reproducible; here the
// mask is a counter-based hash of (seed, call, element index): reproducible, order-independent, no generator state.
__device__ __forceinline__ uint32_t mix_u64(uint64_t z) {
  z = (z ^ (z >> 33)) * 0xff51afd7ed558ccdULL;
  z = (z ^ (z >> 33)) * 0xc4ceb9fe1a85ec53ULL;
  return (uint32_t)((z ^ (z >> 33)) >> 32);
}
__global__ void dropout_kernel(size_t n, float keep, float scale, uint64_t seed, uint64_t call, const float* __restrict__ in,
                               uint8_t* __restrict__ mask, float* __restrict__ out) {
  const uint64_t base = seed * 0x9E3779B97F4A7C15ULL + call * 0xD1B54A32D192ED03ULL;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float u = (float)(mix_u64(base + (uint64_t)i * 0x2545F4914F6CDD1DULL) >> 8) * (1.0f / 16777216.0f);  // [0, 1)
    const uint8_t m = u < keep ? 1 : 0;
    mask[i] = m;
    out[i] = __fmul_rn(__fmul_rn(in[i], (float)m), scale);
  }
}
__global__ void d_dropout_kernel(size_t n, float scale, const float* __restrict__ in, const uint8_t* __restrict__ mask, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __fmul_rn(__fmul_rn(in[i], (float)mask[i]), scale);
}

__global__ void fill_kernel(size_t n, float value, float* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = value;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp per row (the reference GPU kernel is thread-per-row with strided, uncoalesced access: math_functions.cu:158-205).
__global__ void l2norm_kernel(int n, int dim, const float* __restrict__ in, size_t ld_in, float* __restrict__ out, size_t ld_out) {
  const int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* x = in + (size_t)row * ld_in;
  float s = 0.f;
  for (int j = lane; j < dim; j += 32) s += x[j] * x[j];
  s = warp_sum(s);
  s = s < 1.0e-12f ? 1.0e-12f : s;  // l2norm_layer.cpp:29
  s = sqrtf(s);
  for (int j = lane; j < dim; j += 32) out[(size_t)row * ld_out + j] = x[j] / s;
}

// l2norm_layer.cpp:45-62: grad_out = x*coef0*coef1 + g*sum_x2*coef1, coef0 = -<x,g>, coef1 = sum_x2^-1.5
__global__ void d_l2norm_kernel(int n, int dim, const float* __restrict__ feat_in, size_t ld_x, const float* __restrict__ grad_in, size_t ld_g,
                                float* __restrict__ grad_out, size_t ld_o) {
  const int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* x = feat_in + (size_t)row * ld_x;
  const float* g = grad_in + (size_t)row * ld_g;
  float sx = 0.f, c0 = 0.f;
  for (int j = lane; j < dim; j += 32) { sx += x[j] * x[j]; c0 -= x[j] * g[j]; }
  sx = warp_sum(sx);
  c0 = warp_sum(c0);
  sx = sx < 1.0e-12f ? 1.0e-12f : sx;
  const float c1 = powf(sx, -1.5f);
  for (int j = lane; j < dim; j += 32) grad_out[(size_t)row * ld_o + j] = x[j] * c0 * c1 + g[j] * sx * c1;
}

// Masked softmax + cross entropy, one warp per row in [begin, end).
__global__ void softmax_ce_fwd_kernel(int ncls, size_t begin, size_t end, const uint8_t* __restrict__ masks, const uint8_t* __restrict__ labels,
                                      const float* __restrict__ logits, size_t ld_logits, float* __restrict__ probs, size_t ld_probs,
                                      float* __restrict__ losses) {
  const size_t row = begin + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= end) return;
  if (masks && masks[row] != 1) return;
  const float* x = logits + row * ld_logits;
  float mx = -INFINITY;
  for (int j = lane; j < ncls; j += 32) mx = fmaxf(mx, x[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < ncls; j += 32) s += expf(x[j] - mx);
  s = warp_sum(s);
  const int lab = labels[row];
  for (int j = lane; j < ncls; j += 32) {
    const float p = expf(x[j] - mx) / s;
    probs[row * ld_probs + j] = p;
    if (j == lab) losses[row] = -((p == 0.f) ? logf(1e-10f) : logf(p));  // math_functions.cpp:531-542
  }
}

__global__ void softmax_ce_bwd_kernel(int ncls, size_t begin, size_t end, const uint8_t* __restrict__ masks, const uint8_t* __restrict__ labels,
                                      const float* __restrict__ probs, size_t ld_probs, float* __restrict__ grad, double denom, size_t ld_grad) {
  const size_t n = (end - begin) * ncls;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const double inv_range = denom;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const size_t row = begin + k / ncls;
    const int j = (int)(k % ncls);
    if (masks && masks[row] != 1) continue;
    // softmax_loss_layer.cpp:31: (pred - onehot) / (end - begin), evaluated in double then rounded
    grad[row * ld_grad + j] = (float)(((double)probs[row * ld_probs + j] - (labels[row] == j ? 1.0 : 0.0)) / inv_range);
  }
}

// Two-stage, fixed-order reduction (deterministic run to run). Stage 1: one warp per row computes argmax(logits) with the
// reference's tie-break (first maximum, math_functions.cpp:129-139); each CTA folds its rows into one partial
// {loss sum, correct, count}. Stage 2: one CTA adds the partials in index order. stats = {mean loss, accuracy, count}.
__global__ void loss_acc_stage1(int ncls, size_t begin, size_t end, const uint8_t* __restrict__ masks, const uint8_t* __restrict__ labels,
                                const float* __restrict__ logits, size_t ld_logits, const float* __restrict__ losses, double* __restrict__ partial) {
  __shared__ double s_l[8];
  __shared__ unsigned s_c[8], s_n[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t nwarps = (size_t)gridDim.x * 8;
  double l = 0.0; unsigned c = 0, n = 0;
  for (size_t row = begin + (size_t)blockIdx.x * 8 + warp; row < end; row += nwarps) {
    if (masks && masks[row] != 1) continue;
    const float* x = logits + row * ld_logits;
    float mx = -INFINITY; int am = 0x7fffffff;
    for (int j = lane; j < ncls; j += 32) { const float v = x[j]; if (v > mx) { mx = v; am = j; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oam = __shfl_xor_sync(0xffffffffu, am, o);
      if (omx > mx || (omx == mx && oam < am)) { mx = omx; am = oam; }
    }
    if (lane == 0) { l += (double)losses[row]; n += 1; c += (am == (int)labels[row]); }
  }
  if (lane == 0) { s_l[warp] = l; s_c[warp] = c; s_n[warp] = n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tl = 0.0; unsigned tc = 0, tn = 0;
    for (int w = 0; w < 8; w++) { tl += s_l[w]; tc += s_c[w]; tn += s_n[w]; }
    partial[(size_t)blockIdx.x * 3 + 0] = tl;
    partial[(size_t)blockIdx.x * 3 + 1] = (double)tc;
    partial[(size_t)blockIdx.x * 3 + 2] = (double)tn;
  }
}
// ---- vectorised loss kernels: a group of G lanes owns one row (G = 4..32 by class count), 128-bit accesses on row-padded layouts ----
// Forward: probs = softmax(logits), losses[row] = -log(probs[label]) (math_functions.cpp:46-74,531-542), and — when `partial` is given —
// the masked_avg_loss / masked_accuracy_single statistics of the same rows (argmax of the LOGITS with the reference's first-maximum
// tie-break, math_functions.cpp:79-92,129-139) folded into one {loss sum, correct, count} partial per CTA, in a fixed order.
template <int G>
__global__ void __launch_bounds__(256) softmax_ce_rows_kernel(int ncls, size_t begin, size_t end, const uint8_t* __restrict__ masks,
                                                              const uint8_t* __restrict__ labels, const float* __restrict__ logits, size_t ld_logits,
                                                              float* __restrict__ probs, size_t ld_probs, float* __restrict__ losses,
                                                              double* __restrict__ partial) {
  constexpr int GPB = 256 / G;
  __shared__ double s_l[256];
  __shared__ unsigned s_c[256], s_n[256];
  const int gid = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  const int nch = (ncls + 3) >> 2;
  const int kk = (nch + G - 1) / G;
  double l = 0.0; unsigned c = 0, n = 0;
  for (size_t row = begin + (size_t)blockIdx.x * GPB + gid; row < end; row += (size_t)gridDim.x * GPB) {
    // the mask byte, the label and the logits are requested together (rows of [begin, end) are always readable): testing the mask
    // first would put two dependent memory round trips on every row
    const unsigned mk = masks ? masks[row] : 1u;
    const int lab = labels[row];
    const float4* x4 = reinterpret_cast<const float4*>(logits + row * ld_logits);
    float4 tq[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      tq[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (k < kk && gl + G * k < nch) tq[k] = x4[gl + G * k];
    }
    if (mk != 1u) continue;  // uniform within the group
    float v[16];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (k < kk) {  // kk = float4 chunks per lane actually needed (warp-uniform): 1 for up to 4*G classes
        const int ch = gl + G * k;
        const float4 t = tq[k];
        const int col = ch * 4;
        v[4 * k + 0] = col + 0 < ncls ? t.x : -INFINITY; v[4 * k + 1] = col + 1 < ncls ? t.y : -INFINITY;
        v[4 * k + 2] = col + 2 < ncls ? t.z : -INFINITY; v[4 * k + 3] = col + 3 < ncls ? t.w : -INFINITY;
      }
    }
    float mx = -INFINITY; int am = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (k < kk) {
#pragma unroll
        for (int i = 0; i < 4; i++) { if (v[4 * k + i] > mx) { mx = v[4 * k + i]; am = (gl + G * k) * 4 + i; } }
      }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const float omx = __shfl_xor_sync(gmask, mx, o);
      const int oam = __shfl_xor_sync(gmask, am, o);
      if (omx > mx || (omx == mx && oam < am)) { mx = omx; am = oam; }
    }
    float e[16], sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (k < kk) {
#pragma unroll
        for (int i = 0; i < 4; i++) { e[4 * k + i] = expf(v[4 * k + i] - mx); sum += e[4 * k + i]; }
      }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int ch = gl + G * k;
      if (k < kk && ch < nch) {
        float4 pr;
        pr.x = e[4 * k + 0] / sum; pr.y = e[4 * k + 1] / sum; pr.z = e[4 * k + 2] / sum; pr.w = e[4 * k + 3] / sum;
        reinterpret_cast<float4*>(probs + row * ld_probs)[ch] = pr;
        if ((lab >> 2) == ch) {
          const float pl = (lab & 3) == 0 ? pr.x : ((lab & 3) == 1 ? pr.y : ((lab & 3) == 2 ? pr.z : pr.w));
          const float ls = -((pl == 0.f) ? logf(1e-10f) : logf(pl));  // math_functions.cpp:531-542
          losses[row] = ls;
          l += (double)ls;
        }
      }
    }
    if (gl == 0) { n += 1; c += (am == lab); }
  }
  if (partial == nullptr) return;
  s_l[threadIdx.x] = l; s_c[threadIdx.x] = c; s_n[threadIdx.x] = n;
  __syncthreads();
  if (threadIdx.x < 32) {  // fixed-order fold: lane t adds entries t, t+32, ... then a fixed shuffle tree
    double tl = 0.0; unsigned tc = 0, tn = 0;
    for (int i = threadIdx.x; i < 256; i += 32) { tl += s_l[i]; tc += s_c[i]; tn += s_n[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tl += __shfl_down_sync(0xffffffffu, tl, o); tc += __shfl_down_sync(0xffffffffu, tc, o); tn += __shfl_down_sync(0xffffffffu, tn, o);
    }
    if (threadIdx.x == 0) {
      partial[(size_t)blockIdx.x * 3 + 0] = tl; partial[(size_t)blockIdx.x * 3 + 1] = (double)tc; partial[(size_t)blockIdx.x * 3 + 2] = (double)tn;
    }
  }
}

// grad = (probs - onehot) / denom, evaluated in double then rounded (softmax_loss_layer.cpp:31), same row grouping
template <int G>
__global__ void __launch_bounds__(256) softmax_ce_bwd_rows_kernel(int ncls, size_t begin, size_t end, const uint8_t* __restrict__ masks,
                                                                  const uint8_t* __restrict__ labels, const float* __restrict__ probs, size_t ld_probs,
                                                                  float* __restrict__ grad, size_t ld_grad, double denom) {
  constexpr int GPB = 256 / G;
  const int gid = threadIdx.x / G, gl = threadIdx.x % G;
  const int nch = (ncls + 3) >> 2;
  for (size_t row = begin + (size_t)blockIdx.x * GPB + gid; row < end; row += (size_t)gridDim.x * GPB) {
    const unsigned mk = masks ? masks[row] : 1u;
    const int lab = labels[row];
    float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gl < nch) p0 = reinterpret_cast<const float4*>(probs + row * ld_probs)[gl];  // requested before the mask is known
    if (mk != 1u) continue;
    for (int ch = gl; ch < nch; ch += G) {
      const float4 pr = ch == gl ? p0 : reinterpret_cast<const float4*>(probs + row * ld_probs)[ch];
      const int col = ch * 4;
      float4 g;
      g.x = (float)(((double)pr.x - (lab == col + 0 ? 1.0 : 0.0)) / denom); g.y = (float)(((double)pr.y - (lab == col + 1 ? 1.0 : 0.0)) / denom);
      g.z = (float)(((double)pr.z - (lab == col + 2 ? 1.0 : 0.0)) / denom); g.w = (float)(((double)pr.w - (lab == col + 3 ? 1.0 : 0.0)) / denom);
      reinterpret_cast<float4*>(grad + row * ld_grad)[ch] = g;
    }
  }
}

// final fold of the per-CTA partials, one warp, fixed order
__global__ void loss_acc_stage2_warp(int nparts, const double* __restrict__ partial, float* __restrict__ stats) {
  double tl = 0.0, tc = 0.0, tn = 0.0;
  for (int p = threadIdx.x; p < nparts; p += 32) { tl += partial[p * 3 + 0]; tc += partial[p * 3 + 1]; tn += partial[p * 3 + 2]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tl += __shfl_down_sync(0xffffffffu, tl, o); tc += __shfl_down_sync(0xffffffffu, tc, o); tn += __shfl_down_sync(0xffffffffu, tn, o);
  }
  if (threadIdx.x == 0) {
    stats[0] = tn > 0.0 ? (float)(tl / tn) : 0.f;
    stats[1] = (float)tc / (float)tn;  // accuracy_all / float(num_samples), math_functions.cpp:91
    stats[2] = (float)tn;
  }
}

// ---- sigmoid (multi-label) loss: sigmoid_loss_layer::forward/backward (src/layers/sigmoid_loss_layer.cpp:4-36), sigmoid and
// sigmoid_cross_entropy with the reference's mixed float/double expressions (math_functions.cpp:517-521,553-559). One warp per row;
// labels are [nv x ncls] multi-hot bytes.
__global__ void sigmoid_ce_fwd_kernel(int ncls, size_t begin, size_t end, const uint8_t* __restrict__ masks, const uint8_t* __restrict__ labels,
                                      const float* __restrict__ logits, size_t ld_logits, float* __restrict__ probs, size_t ld_probs,
                                      float* __restrict__ losses) {
  const size_t row = begin + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= end) return;
  if (masks && masks[row] != 1) return;
  const float* x = logits + row * ld_logits;
  const uint8_t* y = labels + row * (size_t)ncls;
  float loss = 0.f;
  for (int j = lane; j < ncls; j += 32) {
    const float p = x[j];
    probs[row * ld_probs + j] = (float)(1.0 / (1.0 + (double)expf(-p)));
    const int pos = p >= 0.f ? 1 : 0;
    const float e = expf((float)((double)p - 2.0 * (double)p * (double)pos));
    loss -= __fsub_rn(__fmul_rn(p, (float)y[j] - (float)pos), logf((float)(1.0 + (double)e)));
  }
  loss = warp_sum(loss);
  if (lane == 0) losses[row] = loss;
}
__global__ void sigmoid_ce_bwd_kernel(int ncls, size_t begin, size_t end, const uint8_t* __restrict__ masks, const uint8_t* __restrict__ labels,
                                      const float* __restrict__ probs, size_t ld_probs, float* __restrict__ grad, size_t ld_grad, float denom) {
  const size_t n = (end - begin) * ncls;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const size_t row = begin + k / ncls;
    const int j = (int)(k % ncls);
    if (masks && masks[row] != 1) continue;
    grad[row * ld_grad + j] = __fdiv_rn(probs[row * ld_probs + j] - (float)labels[row * (size_t)ncls + j], denom);  // sigmoid_loss_layer.cpp:29
  }
}
// mean of losses over the masked rows (sigmoid_loss_layer::get_prediction_loss): per-CTA double partials, fixed-order fold
__global__ void loss_mean_stage1(size_t begin, size_t end, const uint8_t* __restrict__ masks, const float* __restrict__ losses,
                                 double* __restrict__ partial) {
  __shared__ double s_l[256];
  __shared__ unsigned s_n[256];
  double l = 0.0; unsigned n = 0;
  for (size_t row = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; row < end; row += (size_t)gridDim.x * blockDim.x)
    if (!masks || masks[row] == 1) { l += (double)losses[row]; n++; }
  s_l[threadIdx.x] = l; s_n[threadIdx.x] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tl = 0.0; unsigned tn = 0;
    for (int i = 0; i < 256; i++) { tl += s_l[i]; tn += s_n[i]; }
    partial[(size_t)blockIdx.x * 3 + 0] = tl; partial[(size_t)blockIdx.x * 3 + 1] = 0.0; partial[(size_t)blockIdx.x * 3 + 2] = (double)tn;
  }
}
// micro-F1 at threshold 0.5 over the (row, class) pairs of the masked rows (masked_f1_score, math_functions.cpp:580-623): integer
// counts (deterministic), then 2PR/(P+R) in double
__global__ void f1_count_kernel(int ncls, size_t begin, size_t end, const uint8_t* __restrict__ masks, const uint8_t* __restrict__ labels,
                                const float* __restrict__ preds, size_t ld_preds, unsigned long long* __restrict__ counts /*tp, fp, fn*/) {
  unsigned tp = 0, fp = 0, fn = 0;
  const size_t n = (end - begin) * ncls;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const size_t row = begin + k / ncls;
    const int j = (int)(k % ncls);
    if (masks && masks[row] != 1) continue;
    const bool pos = preds[row * ld_preds + j] > 0.5f;
    const uint8_t y = labels[row * (size_t)ncls + j];
    tp += (y == 1 && pos); fp += (y == 0 && pos); fn += (y == 1 && !pos);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, o); fp += __shfl_xor_sync(0xffffffffu, fp, o); fn += __shfl_xor_sync(0xffffffffu, fn, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(counts + 0, (unsigned long long)tp); atomicAdd(counts + 1, (unsigned long long)fp); atomicAdd(counts + 2, (unsigned long long)fn); }
}
__global__ void f1_final_kernel(const unsigned long long* __restrict__ counts, float* __restrict__ f1) {
  const double tp = (double)counts[0], fp = (double)counts[1], fn = (double)counts[2];
  const double prec = tp + fp > 0 ? tp / (tp + fp) : 0.0, rec = tp + fn > 0 ? tp / (tp + fn) : 0.0;
  *f1 = (float)(rec + prec > 0.0 ? 2.0 * (rec * prec) / (rec + prec) : 0.0);
}

// optimizer.cpp:22-35 — eps inside the sqrt; b1_t/b2_t are the powers BEFORE this call's post-multiply.
__global__ void adam_kernel(size_t n, const float* __restrict__ dW, float* __restrict__ W, float* __restrict__ m, float* __restrict__ v,
                            float lr, float b1, float b2, float b1_t, float b2_t, float eps) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = dW[i];
    const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(__fsub_rn(1.0f, b1), g));
    const float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(__fmul_rn(__fsub_rn(1.0f, b2), g), g));
    m[i] = mi; v[i] = vi;
    const float num = __fmul_rn(lr, __fdiv_rn(mi, __fsub_rn(1.0f, b1_t)));
    const float den = __fsqrt_rn(__fadd_rn(__fdiv_rn(vi, __fsub_rn(1.0f, b2_t)), eps));
    W[i] = __fsub_rn(W[i], __fdiv_rn(num, den));
  }
}

__global__ void gather_rows_kernel(size_t n_ids, const uint32_t* __restrict__ ids, int F, const float* __restrict__ src, int ld_src,
                                   float* __restrict__ dst, int ld_dst, int vec) {
  // one warp per row; 128-bit loads when the layout allows (src may be a peer-mapped NVLink pointer)
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp; r < n_ids; r += nwarps) {
    const float* s = src + (size_t)ids[r] * ld_src;
    float* d = dst + r * ld_dst;
    if (vec == 4) {
      for (int j = lane; j < F / 4; j += 32) reinterpret_cast<float4*>(d)[j] = reinterpret_cast<const float4*>(s)[j];
    } else {
      for (int j = lane; j < F; j += 32) d[j] = s[j];
    }
  }
}

}  // namespace

extern "C" {

int gai_relu(size_t n, const float* in, float* out, gai_stream_t stream) {
  if (n == 0) return GAI_OK;
  GAI_CHECK_ARG(in && out);
  relu_kernel<<<grid_for(n, 256, 4), 256, 0, gai::S(stream)>>>(n, in, out);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_d_relu(size_t n, const float* grad, const float* data, float* out, gai_stream_t stream) {
  if (n == 0) return GAI_OK;
  GAI_CHECK_ARG(grad && data && out);
  d_relu_kernel<<<grid_for(n, 256, 4), 256, 0, gai::S(stream)>>>(n, grad, data, out);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_d_relu_ld(size_t rows, int F, const float* grad, size_t ld_grad, const float* data, size_t ld_data, float* out, size_t ld_out,
                  gai_stream_t stream) {
  if (rows == 0 || F <= 0) return GAI_OK;
  GAI_CHECK_ARG(grad && data && out && ld_grad >= (size_t)F && ld_data >= (size_t)F && ld_out >= (size_t)F);
  if (ld_grad == (size_t)F && ld_data == (size_t)F && ld_out == (size_t)F) return gai_d_relu(rows * F, grad, data, out, stream);
  d_relu_ld_kernel<<<grid_for(rows * F, 256), 256, 0, gai::S(stream)>>>(rows, F, grad, ld_grad, data, ld_data, out, ld_out);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_dropout(size_t n, float rate, float scale, uint64_t seed, uint64_t call, const float* in, uint8_t* mask, float* out, gai_stream_t stream) {
  if (n == 0) return GAI_OK;
  GAI_CHECK_ARG(in && mask && out && rate >= 0.f && rate < 1.f);
  dropout_kernel<<<grid_for(n, 256, 4), 256, 0, gai::S(stream)>>>(n, 1.0f - rate, scale, seed, call, in, mask, out);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_d_dropout(size_t n, float scale, const float* in, const uint8_t* mask, float* out, gai_stream_t stream) {
  if (n == 0) return GAI_OK;
  GAI_CHECK_ARG(in && mask && out);
  d_dropout_kernel<<<grid_for(n, 256, 4), 256, 0, gai::S(stream)>>>(n, scale, in, mask, out);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_fill(size_t n, float value, float* out, gai_stream_t stream) {
  if (n == 0) return GAI_OK;
  GAI_CHECK_ARG(out != nullptr);
  fill_kernel<<<grid_for(n, 256, 4), 256, 0, gai::S(stream)>>>(n, value, out);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_l2norm_ld(int n, int dim, const float* in, size_t ld_in, float* out, size_t ld_out, gai_stream_t stream) {
  if (n <= 0) return GAI_OK;
  GAI_CHECK_ARG(in && out && dim > 0 && ld_in >= (size_t)dim && ld_out >= (size_t)dim);
  l2norm_kernel<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, gai::S(stream)>>>(n, dim, in, ld_in, out, ld_out);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_l2norm(int n, int dim, const float* in, float* out, gai_stream_t stream) { return gai_l2norm_ld(n, dim, in, (size_t)dim, out, (size_t)dim, stream); }
int gai_d_l2norm_ld(int n, int dim, const float* feat_in, size_t ld_feat, const float* grad_in, size_t ld_grad_in, float* grad_out,
                    size_t ld_grad_out, gai_stream_t stream) {
  if (n <= 0) return GAI_OK;
  GAI_CHECK_ARG(feat_in && grad_in && grad_out && dim > 0 && ld_feat >= (size_t)dim && ld_grad_in >= (size_t)dim && ld_grad_out >= (size_t)dim);
  d_l2norm_kernel<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, gai::S(stream)>>>(n, dim, feat_in, ld_feat, grad_in, ld_grad_in, grad_out,
                                                                                         ld_grad_out);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_d_l2norm(int n, int dim, const float* feat_in, const float* grad_in, float* grad_out, gai_stream_t stream) {
  return gai_d_l2norm_ld(n, dim, feat_in, (size_t)dim, grad_in, (size_t)dim, grad_out, (size_t)dim, stream);
}

// rows are 128-bit addressable (pitch and base multiples of 4 floats, room for the tail group) and fit 4 float4 per lane
static bool loss_vec_ok(int ncls, const float* a, size_t lda, const float* b, size_t ldb) {
  const size_t need = (size_t)((ncls + 3) / 4) * 4;
  return ncls <= 512 && lda % 4 == 0 && ldb % 4 == 0 && lda >= need && ldb >= need && reinterpret_cast<uintptr_t>(a) % 16 == 0 &&
         reinterpret_cast<uintptr_t>(b) % 16 == 0;
}
static int loss_group(int ncls) {  // lanes per row: one float4 per lane up to 32 lanes (128 classes), then up to 4 per lane
  const int nch = (ncls + 3) / 4;
  int G = 4;
  while (G < 32 && G < nch) G <<= 1;
  return G;
}
static int softmax_ce_forward_impl(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                                   size_t ld_logits, float* probs, size_t ld_probs, float* losses, float* stats_d, gai_stream_t stream) {
  GAI_CHECK_ARG(ncls > 0 && begin <= end && labels && logits && probs && losses && ld_logits >= (size_t)ncls && ld_probs >= (size_t)ncls);
  cudaStream_t st = gai::S(stream);
  const size_t rows = end - begin;
  if (rows && loss_vec_ok(ncls, logits, ld_logits, probs, ld_probs)) {
    const int G = loss_group(ncls);
    size_t blocks = (rows * G + 255) / 256;
    const size_t cap = (size_t)gai::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    double* partial = nullptr;
    if (stats_d) {
      void* ws = nullptr;
      int rc = gai::workspace(sizeof(double) * 3 * blocks, &ws, gai::S(stream));
      if (rc != GAI_OK) return rc;
      partial = reinterpret_cast<double*>(ws);
    }
    const unsigned gb = (unsigned)blocks;
    if (G == 4) softmax_ce_rows_kernel<4><<<gb, 256, 0, st>>>(ncls, begin, end, masks, labels, logits, ld_logits, probs, ld_probs, losses, partial);
    else if (G == 8) softmax_ce_rows_kernel<8><<<gb, 256, 0, st>>>(ncls, begin, end, masks, labels, logits, ld_logits, probs, ld_probs, losses, partial);
    else if (G == 16) softmax_ce_rows_kernel<16><<<gb, 256, 0, st>>>(ncls, begin, end, masks, labels, logits, ld_logits, probs, ld_probs, losses, partial);
    else softmax_ce_rows_kernel<32><<<gb, 256, 0, st>>>(ncls, begin, end, masks, labels, logits, ld_logits, probs, ld_probs, losses, partial);
    GAI_LAUNCH_CHECK();
    if (stats_d) {
      loss_acc_stage2_warp<<<1, 32, 0, st>>>((int)blocks, partial, stats_d);
      GAI_LAUNCH_CHECK();
    }
    return GAI_OK;
  }
  if (rows) {
    softmax_ce_fwd_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(ncls, begin, end, masks, labels, logits, ld_logits, probs, ld_probs, losses);
    GAI_LAUNCH_CHECK();
  }
  if (stats_d) return gai_masked_loss_accuracy_ld(ncls, begin, end, masks, labels, logits, ld_logits, losses, stats_d, stream);
  return GAI_OK;
}
int gai_softmax_ce_forward_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                              size_t ld_logits, float* probs, size_t ld_probs, float* losses, gai_stream_t stream) {
  return softmax_ce_forward_impl(ncls, begin, end, masks, labels, logits, ld_logits, probs, ld_probs, losses, nullptr, stream);
}
int gai_softmax_ce_forward_stats_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                                    size_t ld_logits, float* probs, size_t ld_probs, float* losses, float* stats_d, gai_stream_t stream) {
  GAI_CHECK_ARG(stats_d != nullptr);
  return softmax_ce_forward_impl(ncls, begin, end, masks, labels, logits, ld_logits, probs, ld_probs, losses, stats_d, stream);
}
int gai_softmax_ce_forward(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits, float* probs,
                           float* losses, gai_stream_t stream) {
  return gai_softmax_ce_forward_ld(ncls, begin, end, masks, labels, logits, (size_t)ncls, probs, (size_t)ncls, losses, stream);
}
int gai_softmax_ce_backward_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* probs,
                               size_t ld_probs, float* grad_out, size_t ld_grad, uint64_t denom, gai_stream_t stream) {
  GAI_CHECK_ARG(ncls > 0 && begin <= end && labels && probs && grad_out && denom > 0 && ld_grad >= (size_t)ncls && ld_probs >= (size_t)ncls);
  if (begin == end) return GAI_OK;
  cudaStream_t st = gai::S(stream);
  if (loss_vec_ok(ncls, probs, ld_probs, grad_out, ld_grad)) {
    const int nch = (ncls + 3) / 4;
    int G = 4;
    while (G < 32 && G < nch) G <<= 1;
    size_t blocks = ((end - begin) * G + 255) / 256;
    const size_t cap = (size_t)gai::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const unsigned gb = (unsigned)blocks;
    if (G == 4) softmax_ce_bwd_rows_kernel<4><<<gb, 256, 0, st>>>(ncls, begin, end, masks, labels, probs, ld_probs, grad_out, ld_grad, (double)denom);
    else if (G == 8) softmax_ce_bwd_rows_kernel<8><<<gb, 256, 0, st>>>(ncls, begin, end, masks, labels, probs, ld_probs, grad_out, ld_grad, (double)denom);
    else if (G == 16) softmax_ce_bwd_rows_kernel<16><<<gb, 256, 0, st>>>(ncls, begin, end, masks, labels, probs, ld_probs, grad_out, ld_grad, (double)denom);
    else softmax_ce_bwd_rows_kernel<32><<<gb, 256, 0, st>>>(ncls, begin, end, masks, labels, probs, ld_probs, grad_out, ld_grad, (double)denom);
    GAI_LAUNCH_CHECK();
    return GAI_OK;
  }
  softmax_ce_bwd_kernel<<<grid_for((end - begin) * ncls, 256), 256, 0, st>>>(ncls, begin, end, masks, labels, probs, ld_probs, grad_out, (double)denom,
                                                                            ld_grad);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_softmax_ce_backward(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* probs, float* grad_out,
                            gai_stream_t stream) {
  if (begin == end) return GAI_OK;
  return gai_softmax_ce_backward_ld(ncls, begin, end, masks, labels, probs, (size_t)ncls, grad_out, (size_t)ncls, (uint64_t)(end - begin), stream);
}
int gai_softmax_ce_backward_scaled(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* probs,
                                   float* grad_out, int ld_grad, uint64_t denom, gai_stream_t stream) {
  GAI_CHECK_ARG(ld_grad >= ncls);
  return gai_softmax_ce_backward_ld(ncls, begin, end, masks, labels, probs, (size_t)ncls, grad_out, (size_t)ld_grad, denom, stream);
}
int gai_masked_loss_accuracy_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                                size_t ld_logits, const float* losses, float* stats_d, gai_stream_t stream) {
  GAI_CHECK_ARG(ncls > 0 && begin <= end && labels && logits && losses && stats_d && ld_logits >= (size_t)ncls);
  size_t rows = end - begin;
  int nparts = (int)((rows + 7) / 8);
  const int cap = gai::sm_count() * 8;
  if (nparts > cap) nparts = cap;
  if (nparts < 1) nparts = 1;
  void* ws = nullptr;
  int rc = gai::workspace(sizeof(double) * 3 * (size_t)nparts, &ws, gai::S(stream));
  if (rc != GAI_OK) return rc;
  loss_acc_stage1<<<nparts, 256, 0, gai::S(stream)>>>(ncls, begin, end, masks, labels, logits, ld_logits, losses, reinterpret_cast<double*>(ws));
  GAI_LAUNCH_CHECK();
  loss_acc_stage2_warp<<<1, 32, 0, gai::S(stream)>>>(nparts, reinterpret_cast<const double*>(ws), stats_d);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_masked_loss_accuracy(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                             const float* losses, float* stats_d, gai_stream_t stream) {
  return gai_masked_loss_accuracy_ld(ncls, begin, end, masks, labels, logits, (size_t)ncls, losses, stats_d, stream);
}

int gai_sigmoid_ce_forward_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels_multi, const float* logits,
                              size_t ld_logits, float* probs, size_t ld_probs, float* losses, gai_stream_t stream) {
  GAI_CHECK_ARG(ncls > 0 && begin <= end && labels_multi && logits && probs && losses && ld_logits >= (size_t)ncls && ld_probs >= (size_t)ncls);
  if (begin == end) return GAI_OK;
  sigmoid_ce_fwd_kernel<<<(unsigned)(((end - begin) * 32 + 255) / 256), 256, 0, gai::S(stream)>>>(ncls, begin, end, masks, labels_multi, logits,
                                                                                                    ld_logits, probs, ld_probs, losses);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_sigmoid_ce_backward_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels_multi, const float* probs,
                               size_t ld_probs, float* grad_out, size_t ld_grad, uint64_t denom, gai_stream_t stream) {
  GAI_CHECK_ARG(ncls > 0 && begin <= end && labels_multi && probs && grad_out && denom > 0 && ld_grad >= (size_t)ncls && ld_probs >= (size_t)ncls);
  if (begin == end) return GAI_OK;
  sigmoid_ce_bwd_kernel<<<grid_for((end - begin) * ncls, 256), 256, 0, gai::S(stream)>>>(ncls, begin, end, masks, labels_multi, probs, ld_probs,
                                                                                         grad_out, ld_grad, (float)denom);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_masked_loss_mean(size_t begin, size_t end, const uint8_t* masks, const float* losses, float* stats_d, gai_stream_t stream) {
  GAI_CHECK_ARG(begin <= end && losses && stats_d);
  size_t blocks = (end - begin + 255) / 256;
  const size_t cap = (size_t)gai::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  void* ws = nullptr;
  int rc = gai::workspace(sizeof(double) * 3 * blocks, &ws, gai::S(stream));
  if (rc != GAI_OK) return rc;
  loss_mean_stage1<<<(unsigned)blocks, 256, 0, gai::S(stream)>>>(begin, end, masks, losses, reinterpret_cast<double*>(ws));
  GAI_LAUNCH_CHECK();
  loss_acc_stage2_warp<<<1, 32, 0, gai::S(stream)>>>((int)blocks, reinterpret_cast<const double*>(ws), stats_d);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
int gai_masked_f1_micro(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels_multi, const float* preds, size_t ld_preds,
                        float* f1_d, gai_stream_t stream) {
  GAI_CHECK_ARG(ncls > 0 && begin <= end && labels_multi && preds && f1_d && ld_preds >= (size_t)ncls);
  void* ws = nullptr;
  int rc = gai::workspace(sizeof(unsigned long long) * 4, &ws, gai::S(stream));
  if (rc != GAI_OK) return rc;
  GAI_CUDA(cudaMemsetAsync(ws, 0, sizeof(unsigned long long) * 4, gai::S(stream)));
  if (begin != end) {
    f1_count_kernel<<<grid_for((end - begin) * ncls, 256), 256, 0, gai::S(stream)>>>(ncls, begin, end, masks, labels_multi, preds, ld_preds,
                                                                                    reinterpret_cast<unsigned long long*>(ws));
    GAI_LAUNCH_CHECK();
  }
  f1_final_kernel<<<1, 1, 0, gai::S(stream)>>>(reinterpret_cast<const unsigned long long*>(ws), f1_d);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

int gai_adam_update(size_t n, const float* dW, float* W, float* m, float* v, float lr, float b1, float b2, float b1_t, float b2_t, float eps,
                    gai_stream_t stream) {
  if (n == 0) return GAI_OK;
  GAI_CHECK_ARG(dW && W && m && v);
  adam_kernel<<<grid_for(n, 256), 256, 0, gai::S(stream)>>>(n, dW, W, m, v, lr, b1, b2, b1_t, b2_t, eps);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

int gai_gather_rows(size_t n_ids, const uint32_t* ids, int F, const float* src, int ld_src, float* dst, int ld_dst, gai_stream_t stream) {
  if (n_ids == 0) return GAI_OK;
  GAI_CHECK_ARG(ids && src && dst && F > 0 && ld_src >= F && ld_dst >= F);
  const bool v4 = (F % 4 == 0) && (ld_src % 4 == 0) && (ld_dst % 4 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
                  (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  gather_rows_kernel<<<grid_for(n_ids * 32, 256), 256, 0, gai::S(stream)>>>(n_ids, ids, F, src, ld_src, dst, ld_dst, v4 ? 4 : 1);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

int gai_set_gemm_mode(int mode) {
  GAI_CHECK_ARG(mode >= 0 && mode <= 3);
  gai::g_gemm_mode = mode;
  return GAI_OK;
}
int gai_get_gemm_mode(void) { return gai::g_gemm_mode; }

int gai_matmul_ld(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int transA,
                  int transB, int accum, int flags, gai_stream_t stream) {
  if (x == 0 || y == 0) return GAI_OK;
  GAI_CHECK_ARG(A && B && C);
  GAI_CHECK_ARG(lda >= (transA ? x : z) && ldb >= (transB ? z : y) && ldc >= y);
  cudaStream_t st = gai::S(stream);
  const int mode = gai::g_gemm_mode;
  if (mode != 1) {
    // op(A) = A^T with a long reduction over the stored rows (dW = X^T·G): the MN-major split-over-rows kernel
    int rc = (transA && !transB) ? gai::gemm_tc_wgrad(x, y, z, A, lda, B, ldb, C, ldc, accum, flags, mode == 3 ? 1 : 3, st)
                                 : gai::gemm_tc(x, y, z, A, lda, B, ldb, C, ldc, transA, transB, accum, flags, mode == 3 ? 1 : 3, st);
    if (rc != GAI_ERR_UNSUPPORTED) return rc;
    if (mode >= 2) return rc;  // an explicit tensor-core request must not silently degrade
  }
  return gai::gemm_simt(x, y, z, A, lda, B, ldb, C, ldc, transA, transB, accum, flags, st);
}

static int sign_bits(size_t rows, int F, const float* data, size_t ldd, uint32_t* bits, size_t ldb, cudaStream_t st) {
  sign_bits_kernel<<<grid_for(rows * ((F + 31) / 32) * 32, 256), 256, 0, st>>>(rows, F, data, ldd, bits, ldb);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}
static int d_relu_bits(size_t rows, int F, float* grad, size_t ldg, const uint32_t* bits, size_t ldb, cudaStream_t st) {
  d_relu_bits_kernel<<<grid_for(rows * F, 256), 256, 0, st>>>(rows, F, grad, ldg, bits, ldb);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

int gai_matmul_kcat(size_t x, size_t y, size_t z1, const float* A1, size_t lda1, const float* B1, size_t ldb1, size_t z2, const float* A2,
                    size_t lda2, const float* B2, size_t ldb2, float* C, size_t ldc, int transB, int flags, const float* mask, size_t ldmask,
                    uint32_t* relu_bits, size_t ld_bits, gai_stream_t stream) {
  if (x == 0 || y == 0) return GAI_OK;
  GAI_CHECK_ARG(A1 && B1 && A2 && B2 && C && z1 > 0 && z2 > 0);
  GAI_CHECK_ARG(lda1 >= z1 && lda2 >= z2 && ldc >= y && ldb1 >= (transB ? z1 : y) && ldb2 >= (transB ? z2 : y));
  GAI_CHECK_ARG((flags & ~(GAI_EPI_RELU | GAI_EPI_MASK | GAI_EPI_PADDED | GAI_EPI_BITMASK)) == 0);
  const bool bitmask = (flags & GAI_EPI_MASK) && (flags & GAI_EPI_BITMASK);
  GAI_CHECK_ARG(!(flags & GAI_EPI_MASK) || (mask && ldmask >= (bitmask ? (y + 31) / 32 : y) && !(flags & GAI_EPI_RELU)));
  GAI_CHECK_ARG(!relu_bits || ((flags & GAI_EPI_RELU) && ld_bits >= (y + 31) / 32));
  cudaStream_t st = gai::S(stream);
  const int mode = gai::g_gemm_mode;
  if (mode != 1) {
    gai::GemmCat q;
    q.M = x; q.nk = 2; q.tb = transB;
    q.A[0] = A1; q.lda[0] = lda1; q.K[0] = z1; q.B[0][0] = B1; q.ldb[0][0] = ldb1;
    q.A[1] = A2; q.lda[1] = lda2; q.K[1] = z2; q.B[1][0] = B2; q.ldb[1][0] = ldb2;
    q.N[0] = y; q.C[0] = C; q.ldc[0] = ldc; q.flags = flags; q.mask = mask; q.ldmask = ldmask; q.bits_out = relu_bits; q.ld_bits = ld_bits;
    int rc = gai::gemm_tc_cat(q, mode == 3 ? 1 : 3, st);
    if (rc != GAI_ERR_UNSUPPORTED || mode >= 2) return rc;
  }
  // shapes the tensor-core kernel declines (few rows, wide outputs): the same sum as two SIMT products + the mask / sign-bit pass
  const int relu = (flags & GAI_EPI_MASK) ? 0 : (flags & GAI_EPI_RELU);
  int rc = gai::gemm_simt(x, y, z1, A1, lda1, B1, ldb1, C, ldc, 0, transB, 0, 0, st);
  if (rc != GAI_OK) return rc;
  rc = gai::gemm_simt(x, y, z2, A2, lda2, B2, ldb2, C, ldc, 0, transB, 1, relu, st);
  if (rc != GAI_OK) return rc;
  if (bitmask) return d_relu_bits(x, (int)y, C, ldc, reinterpret_cast<const uint32_t*>(mask), ldmask, st);
  if (flags & GAI_EPI_MASK) return gai_d_relu_ld(x, (int)y, C, ldc, mask, ldmask, C, ldc, stream);
  if (relu_bits) return sign_bits(x, (int)y, C, ldc, relu_bits, ld_bits, st);
  return GAI_OK;
}

int gai_matmul_mask(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int transB,
                    const float* mask, size_t ldmask, int flags, gai_stream_t stream) {
  if (x == 0 || y == 0) return GAI_OK;
  const bool bitmask = (flags & GAI_EPI_BITMASK) != 0;
  GAI_CHECK_ARG(A && B && C && mask && z > 0 && lda >= z && ldc >= y && ldmask >= (bitmask ? (y + 31) / 32 : y) && ldb >= (transB ? z : y) &&
                (flags & ~(GAI_EPI_PADDED | GAI_EPI_BITMASK)) == 0);
  cudaStream_t st = gai::S(stream);
  const int mode = gai::g_gemm_mode;
  if (mode != 1) {
    gai::GemmCat q;
    q.M = x; q.tb = transB;
    q.A[0] = A; q.lda[0] = lda; q.K[0] = z; q.B[0][0] = B; q.ldb[0][0] = ldb;
    q.N[0] = y; q.C[0] = C; q.ldc[0] = ldc; q.flags = GAI_EPI_MASK | flags; q.mask = mask; q.ldmask = ldmask;
    int rc = gai::gemm_tc_cat(q, mode == 3 ? 1 : 3, st);
    if (rc != GAI_ERR_UNSUPPORTED || mode >= 2) return rc;
  }
  int rc = gai::gemm_simt(x, y, z, A, lda, B, ldb, C, ldc, 0, transB, 0, 0, st);
  if (rc != GAI_OK) return rc;
  if (bitmask) return d_relu_bits(x, (int)y, C, ldc, reinterpret_cast<const uint32_t*>(mask), ldmask, st);
  return gai_d_relu_ld(x, (int)y, C, ldc, mask, ldmask, C, ldc, stream);
}

// C = ReLU(A·B) and the sign bits of C in one pass (aggregate-first GCN forward, gcn_layer.cpp:20-24 + the mask the layer above needs)
int gai_matmul_relu_bits(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int flags,
                         uint32_t* relu_bits, size_t ld_bits, gai_stream_t stream) {
  if (x == 0 || y == 0) return GAI_OK;
  GAI_CHECK_ARG(A && B && C && relu_bits && z > 0 && lda >= z && ldc >= y && ldb >= y && ld_bits >= (y + 31) / 32 && (flags & ~GAI_EPI_PADDED) == 0);
  cudaStream_t st = gai::S(stream);
  const int mode = gai::g_gemm_mode;
  if (mode != 1) {
    gai::GemmCat q;
    q.M = x;
    q.A[0] = A; q.lda[0] = lda; q.K[0] = z; q.B[0][0] = B; q.ldb[0][0] = ldb;
    q.N[0] = y; q.C[0] = C; q.ldc[0] = ldc; q.flags = GAI_EPI_RELU | flags; q.bits_out = relu_bits; q.ld_bits = ld_bits;
    int rc = gai::gemm_tc_cat(q, mode == 3 ? 1 : 3, st);
    if (rc != GAI_ERR_UNSUPPORTED || mode >= 2) return rc;
  }
  int rc = gai::gemm_simt(x, y, z, A, lda, B, ldb, C, ldc, 0, 0, 0, GAI_EPI_RELU, st);
  if (rc != GAI_OK) return rc;
  return sign_bits(x, (int)y, C, ldc, relu_bits, ld_bits, st);
}

int gai_matmul_ncat(size_t x, size_t z, const float* A, size_t lda, size_t y1, const float* B1, size_t ldb1, float* C1, size_t ldc1, size_t y2,
                    const float* B2, size_t ldb2, float* C2, size_t ldc2, int flags, gai_stream_t stream) {
  if (x == 0) return GAI_OK;
  GAI_CHECK_ARG(A && B1 && B2 && C1 && C2 && z > 0 && y1 > 0 && y2 > 0 && (flags & ~GAI_EPI_PADDED) == 0);
  GAI_CHECK_ARG(lda >= z && ldb1 >= y1 && ldb2 >= y2 && ldc1 >= y1 && ldc2 >= y2);
  cudaStream_t st = gai::S(stream);
  const int mode = gai::g_gemm_mode;
  if (mode != 1) {
    gai::GemmCat q;
    q.M = x; q.nn = 2;
    q.A[0] = A; q.lda[0] = lda; q.K[0] = z;
    q.B[0][0] = B1; q.ldb[0][0] = ldb1; q.N[0] = y1; q.C[0] = C1; q.ldc[0] = ldc1;
    q.B[0][1] = B2; q.ldb[0][1] = ldb2; q.N[1] = y2; q.C[1] = C2; q.ldc[1] = ldc2; q.flags = flags;
    int rc = gai::gemm_tc_cat(q, mode == 3 ? 1 : 3, st);
    if (rc != GAI_ERR_UNSUPPORTED || mode >= 2) return rc;
  }
  int rc = gai::gemm_simt(x, y1, z, A, lda, B1, ldb1, C1, ldc1, 0, 0, 0, 0, st);
  if (rc != GAI_OK) return rc;
  return gai::gemm_simt(x, y2, z, A, lda, B2, ldb2, C2, ldc2, 0, 0, 0, 0, st);
}

static int wgrad_cat_or_simt(const gai::WgradCat& q, cudaStream_t st) {
  const int mode = gai::g_gemm_mode;
  if (mode != 1) {
    int rc = gai::gemm_tc_wgrad_cat(q, mode == 3 ? 1 : 3, st);
    if (rc != GAI_ERR_UNSUPPORTED) return rc;
  }
  // widths the two-operand kernel declines (more than 512 accumulator columns in all): two plain products, each on the tensor
  // cores when its own shape fits, otherwise SIMT
  for (int i = 0; i < 2; i++) {
    const int a = q.dual == 1 ? i : 0, b = q.dual == 2 ? i : 0;
    int rc = GAI_ERR_UNSUPPORTED;
    if (mode != 1) rc = gai::gemm_tc_wgrad(q.Kx[a], q.My[b], q.nrows, q.A[a], q.lda[a], q.B[b], q.ldb[b], q.C[i], q.ldc[i], 0, 0, mode == 3 ? 1 : 3, st);
    if (rc == GAI_ERR_UNSUPPORTED && mode < 2) rc = gai::gemm_simt(q.Kx[a], q.My[b], q.nrows, q.A[a], q.lda[a], q.B[b], q.ldb[b], q.C[i], q.ldc[i], 1, 0, 0, 0, st);
    if (rc != GAI_OK) return rc;
  }
  return GAI_OK;
}

int gai_wgrad_two_a(size_t z, size_t y, const float* B, size_t ldb, size_t x1, const float* A1, size_t lda1, float* C1, size_t ldc1, size_t x2,
                    const float* A2, size_t lda2, float* C2, size_t ldc2, gai_stream_t stream) {
  GAI_CHECK_ARG(B && A1 && A2 && C1 && C2 && z > 0 && y > 0 && x1 > 0 && x2 > 0);
  GAI_CHECK_ARG(ldb >= y && lda1 >= x1 && lda2 >= x2 && ldc1 >= y && ldc2 >= y);
  gai::WgradCat q;
  q.nrows = z; q.dual = 1;
  q.A[0] = A1; q.lda[0] = lda1; q.Kx[0] = x1; q.A[1] = A2; q.lda[1] = lda2; q.Kx[1] = x2;
  q.B[0] = B; q.ldb[0] = ldb; q.My[0] = y;
  q.C[0] = C1; q.ldc[0] = ldc1; q.C[1] = C2; q.ldc[1] = ldc2;
  return wgrad_cat_or_simt(q, gai::S(stream));
}

int gai_wgrad_two_b(size_t z, size_t x, const float* A, size_t lda, size_t y1, const float* B1, size_t ldb1, float* C1, size_t ldc1, size_t y2,
                    const float* B2, size_t ldb2, float* C2, size_t ldc2, gai_stream_t stream) {
  GAI_CHECK_ARG(A && B1 && B2 && C1 && C2 && z > 0 && x > 0 && y1 > 0 && y2 > 0);
  GAI_CHECK_ARG(lda >= x && ldb1 >= y1 && ldb2 >= y2 && ldc1 >= y1 && ldc2 >= y2);
  gai::WgradCat q;
  q.nrows = z; q.dual = 2;
  q.A[0] = A; q.lda[0] = lda; q.Kx[0] = x;
  q.B[0] = B1; q.ldb[0] = ldb1; q.My[0] = y1; q.B[1] = B2; q.ldb[1] = ldb2; q.My[1] = y2;
  q.C[0] = C1; q.ldc[0] = ldc1; q.C[1] = C2; q.ldc[1] = ldc2;
  return wgrad_cat_or_simt(q, gai::S(stream));
}

int gai_matmul(size_t x, size_t y, size_t z, const float* A, const float* B, float* C, int transA, int transB, int accum, int flags,
               gai_stream_t stream) {
  // reference layout (math_functions.cpp:142-151): lda = transA ? x : z, ldb = transB ? z : y, ldc = y
  return gai_matmul_ld(x, y, z, A, transA ? x : z, B, transB ? z : y, C, y, transA, transB, accum, flags, stream);
}

}  // extern "C"
