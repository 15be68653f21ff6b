// fp32 SIMT (FFMA) dense transform: the exact-fp32 path behind gai_matmul (mode 1) and the fallback for shapes the
// tcgen05 kernel does not take (tiny problems such as cora's 2708x1433x16, where launch latency dominates).
// Replaces sgemm_gpu/cublasSgemm (src/utilities/math_functions.cu:321-343).  Row-major, any transposition, any
// leading dimension, split-K over the reduction with a deterministic in-order second stage (dW = X^T·G has a
// K x M output and an N-long reduction, SURVEY.md §7 hard part 3).
#include "gai_internal.cuh"

namespace gai {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, PAD = 4;

struct GemmArgs {
  const float* A; const float* B; float* C;
  size_t M, N, K;
  size_t lda, ldb, ldc;
  size_t tiles_n;
  size_t k_chunk;      // reduction elements per split
  float* partial;      // [splits][M][N] when splits > 1
  int accum, flags;
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
  __shared__ float As[2][BK][BM + PAD];
  __shared__ float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  // 1-D tile index with the n tile fastest: CTAs that share an A row-panel are adjacent in launch order (L2 reuse).
  const size_t m0 = ((size_t)blockIdx.x / g.tiles_n) * BM, n0 = ((size_t)blockIdx.x % g.tiles_n) * BN;
  const size_t kb = (size_t)blockIdx.y * g.k_chunk;
  const size_t ke = (kb + g.k_chunk < g.K) ? kb + g.k_chunk : g.K;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;

  float ra[8], rb[4];
  auto load_tile = [&](size_t k0) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int e = tid + i * 256;
      int m, k;
      if (TA) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
      const size_t gm = m0 + m, gk = k0 + k;
      float v = 0.0f;
      if (gm < g.M && gk < ke) v = TA ? __ldg(g.A + gk * g.lda + gm) : __ldg(g.A + gm * g.lda + gk);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int e = tid + i * 256;
      int n, k;
      if (TB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      const size_t gn = n0 + n, gk = k0 + k;
      float v = 0.0f;
      if (gn < g.N && gk < ke) v = TB ? __ldg(g.B + gn * g.ldb + gk) : __ldg(g.B + gk * g.ldb + gn);
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int e = tid + i * 256;
      int m, k;
      if (TA) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
      As[buf][k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int e = tid + i * 256;
      int n, k;
      if (TB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      Bs[buf][k][n] = rb[i];
    }
  };

  if (kb < ke) {
    load_tile(kb);
    store_tile(0);
    __syncthreads();
    int buf = 0;
    for (size_t k0 = kb; k0 < ke; k0 += BK) {
      const bool more = (k0 + BK) < ke;
      if (more) load_tile(k0 + BK);
#pragma unroll
      for (int k = 0; k < BK; k++) {
        float a[TM], b[TN];
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
          for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (more) {
        store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  const bool split = gridDim.y > 1;
#pragma unroll
  for (int i = 0; i < TM; i++) {
    const size_t gm = m0 + ty * TM + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const size_t gn = n0 + tx * TN + j;
      if (gn >= g.N) continue;
      if (split) {
        g.partial[((size_t)blockIdx.y * g.M + gm) * g.N + gn] = acc[i][j];
      } else {
        float r = acc[i][j];
        if (g.accum) r += g.C[gm * g.ldc + gn];
        if (g.flags & GAI_EPI_RELU) r = r > 0.0f ? r : 0.0f;
        g.C[gm * g.ldc + gn] = r;
      }
    }
  }
}

// C = (accum ? C : 0) + sum_s partial[s], s ascending (deterministic).
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, size_t M, size_t N, size_t ldc,
                                     int splits, int accum, int flags) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const size_t m = i / N, n = i % N;
  float r = accum ? C[m * ldc + n] : 0.0f;
  for (int s = 0; s < splits; s++) r += partial[(size_t)s * M * N + i];
  if (flags & GAI_EPI_RELU) r = r > 0.0f ? r : 0.0f;
  C[m * ldc + n] = r;
}

int gemm_simt(size_t M, size_t N, size_t K, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int ta, int tb,
              int accum, int flags, cudaStream_t st) {
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.accum = accum; g.flags = flags;
  const size_t tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const size_t tiles = tiles_m * tiles_n;
  const int sms = sm_count();
  int splits = 1;
  if (tiles < (size_t)2 * sms && K >= 512) {
    size_t want = ((size_t)4 * sms + tiles - 1) / tiles;
    size_t maxs = (K + 255) / 256;
    splits = (int)(want < maxs ? want : maxs);
    if (splits < 1) splits = 1;
    if (splits > 2048) splits = 2048;
  }
  size_t k_chunk = (K + splits - 1) / splits;
  k_chunk = ((k_chunk + BK - 1) / BK) * BK;
  splits = (int)((K + k_chunk - 1) / k_chunk);
  if (splits < 1) splits = 1;
  g.k_chunk = k_chunk;
  g.partial = nullptr;
  if (splits > 1) {
    void* ws = nullptr;
    int rc = workspace(sizeof(float) * (size_t)splits * M * N, &ws, st);
    if (rc != GAI_OK) return rc;
    g.partial = reinterpret_cast<float*>(ws);
  }
  g.tiles_n = tiles_n;
  if (tiles >= (size_t(1) << 31)) return set_error(GAI_ERR_UNSUPPORTED, "gemm_simt", "too many tiles");
  dim3 grid((unsigned)tiles, (unsigned)splits, 1);
  if (ta && tb) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
  else if (ta) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (tb) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
  GAI_LAUNCH_CHECK();
  if (splits > 1) {
    const size_t n = M * N;
    splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g.partial, C, M, N, ldc, splits, accum, flags);
    GAI_LAUNCH_CHECK();
  }
  return GAI_OK;
}

}  // namespace gai
