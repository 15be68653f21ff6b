// Dense transform on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with TMEM accumulators, operands staged by
// TMA, fp32 in / fp32 out with 3xTF32 error compensation (A·B ≈ A_lo·B_hi + A_hi·B_lo + A_hi·B_hi, every term accumulated
// in fp32 inside TMEM), so that results stay within the fp32 tolerance of the reference's cblas/cublas SGEMM
// (src/utilities/math_functions.cpp:142-171, math_functions.cu:321-343).
//
// Shapes served (row-major, the tall operand is the N x K activation / gradient matrix, the weight is tiny):
//   C[M x N] = A[M x K] · B[K x N]          (forward transforms X·W)           transB = 0
//   C[M x N] = A[M x K] · B[N x K]^T        (input gradients G·W^T)            transB = 1
// with N <= 256 (one MMA tile covers the full output width) and any M. The reduction-over-rows product X^T·G (weight
// gradient) and anything else is declined (GAI_ERR_UNSUPPORTED) and served by gemm_simt.cu.
//
// Kernel (persistent, one CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: per k-block (32 fp32 = one 128-byte swizzle row) loads the A tile [128 x 32] and the
//               pre-split weight tiles B_hi/B_lo [N x 32] into a multi-stage shared-memory ring (mbarrier expect_tx).
//   warps 8-11  splitter: rewrites the landed A tile in place as A_hi = rn_tf32(A) and writes A_lo = rn_tf32(A - A_hi)
//               next to it (same swizzled offsets), then fence.proxy.async + arrive.
//   warp 1      MMA issuer: one thread issues 3 x 4 tcgen05.mma (M128 x N x K8) per k-block into one of two TMEM
//               accumulators; tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns per warp, optional "+C" and ReLU, 128-byte row stores.
// The weight is prepared once per call by a tiny kernel (transpose to K-major if needed, zero-pad to [Npad x Kpad],
// split into tf32 hi/lo) so that both operands are K-major and TMA-addressable whatever the caller's layout.
#include "tc_common.cuh"

namespace gai {

namespace {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int BK = 32;           // fp32 per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int THREADS = 384;

using namespace tc;

struct TcArgs {
  float* C;
  size_t M, N, ldc;
  int n_mma;       // N rounded up to a multiple of 16 (UMMA N)
  int num_kb;      // k-blocks of 32
  int stages;
  int passes;      // 3 = 3xTF32, 1 = single TF32 pass
  int accum, flags;
  uint32_t stage_bytes, b_tile_bytes;
};

__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
               const __grid_constant__ CUtensorMap map_blo, const TcArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], conv_bar[8], empty_bar[8], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t num_tiles = (g.M + BM - 1) / BM;
  constexpr uint32_t A_BYTES = BM * BK * 4;  // 16 KB

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; i++) { mbar_init(&full_bar[i], 1); mbar_init(&conv_bar[i], 4); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 512 columns = two fp32 accumulators of up to 256 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t it = 0;
      for (size_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < g.num_kb; kb++, it++) {
          const int s = it % g.stages;
          mbar_wait(&empty_bar[s], ((it / g.stages) & 1) ^ 1);
          uint8_t* st = smem + (size_t)s * g.stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s], A_BYTES + (g.passes == 3 ? 2 : 1) * g.b_tile_bytes);
          tma_load_2d(st, &map_a, kb * BK, (int)(tile * BM), &full_bar[s]);
          tma_load_2d(st + 2 * A_BYTES, &map_bhi, kb * BK, 0, &full_bar[s]);
          if (g.passes == 3) tma_load_2d(st + 2 * A_BYTES + g.b_tile_bytes, &map_blo, kb * BK, 0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
      // N >> 3 in bits [17,23), M >> 4 in bits [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.n_mma >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t it = 0, tcount = 0;
      for (size_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
        const int acc = tcount & 1;
        mbar_wait(&tempty_bar[acc], ((tcount >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < g.num_kb; kb++, it++) {
          const int s = it % g.stages;
          mbar_wait(&conv_bar[s], (it / g.stages) & 1);
          tcgen05_fence_after();
          const uint32_t a_hi = smem_u32(smem + (size_t)s * g.stage_bytes);
          const uint32_t a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES;
          const uint32_t b_lo = b_hi + g.b_tile_bytes;
#pragma unroll
          for (int k = 0; k < BK / 8; k++) {
            const uint32_t koff = k * 32;  // 8 tf32 = 32 bytes along the swizzled 128-byte row
            const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
            if (g.passes == 3) {
              umma_tf32(d_tmem, make_desc_k128(a_lo + koff), make_desc_k128(b_hi + koff), idesc, first);
              umma_tf32(d_tmem, make_desc_k128(a_hi + koff), make_desc_k128(b_lo + koff), idesc, 1u);
              umma_tf32(d_tmem, make_desc_k128(a_hi + koff), make_desc_k128(b_hi + koff), idesc, 1u);
            } else {
              umma_tf32(d_tmem, make_desc_k128(a_hi + koff), make_desc_k128(b_hi + koff), idesc, first);
            }
          }
          umma_commit(&empty_bar[s]);  // smem stage reusable once these MMAs have read it
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
      }
    }
  } else if (warp >= 8) {
    // ---------------- splitter: A -> (A_hi in place, A_lo) ----------------
    const int t = threadIdx.x - 256;  // 0..127
    uint32_t it = 0;
    for (size_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < g.num_kb; kb++, it++) {
        const int s = it % g.stages;
        mbar_wait(&full_bar[s], (it / g.stages) & 1);
        if (g.passes == 3) {
          uint4* hi = reinterpret_cast<uint4*>(smem + (size_t)s * g.stage_bytes);
          uint4* lo = reinterpret_cast<uint4*>(smem + (size_t)s * g.stage_bytes + A_BYTES);
#pragma unroll
          for (int i = 0; i < (int)(A_BYTES / 16 / 128); i++) {
            const int idx = t + i * 128;
            uint4 v = hi[idx];
            uint4 h, l;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            hi[idx] = h;
            lo[idx] = l;
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[s]);
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    const int q = warp - 4;  // TMEM lane quarter == warp index within the warpgroup
    uint32_t tcount = 0;
    for (size_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
      const int acc = tcount & 1;
      mbar_wait(&tfull_bar[acc], (tcount >> 1) & 1);
      tcgen05_fence_after();
      const size_t row = tile * BM + (size_t)q * 32 + lane;
      const bool row_ok = row < g.M;
      float* crow = g.C + row * g.ldc;
      const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
      for (int c0 = 0; c0 < g.n_mma; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)c0, r);
        if (row_ok) {
          if (vec_ok && c0 + 32 <= (int)g.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              if (g.accum) { const float4 o = *reinterpret_cast<const float4*>(crow + c0 + j); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
              if (g.flags & GAI_EPI_RELU) { v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f; v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f; }
              *reinterpret_cast<float4*>(crow + c0 + j) = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
              if (c0 + j < (int)g.N) {
                float v = __uint_as_float(r[j]);
                if (g.accum) v += crow[c0 + j];
                if (g.flags & GAI_EPI_RELU) v = v > 0.f ? v : 0.f;
                crow[c0 + j] = v;
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Weight preparation: Bt[n][k] = op(B)[k][n] zero-padded to [n_pad x k_pad], split into tf32 hi / lo (both K-major).
__global__ void prep_b_kernel(const float* __restrict__ B, size_t ldb, int tb, size_t K, size_t N, int k_pad, int n_pad,
                              float* __restrict__ hi, float* __restrict__ lo) {
  const size_t total = (size_t)k_pad * n_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / k_pad, k = i % k_pad;
    float v = 0.f;
    if (n < N && k < K) v = tb ? B[n * ldb + k] : B[k * ldb + n];
    uint32_t h, l;
    split_tf32(__float_as_uint(v), h, l);
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(l);
  }
}

__global__ void pad_a_kernel(size_t M, size_t K, size_t Kp, const float* __restrict__ A, size_t lda, float* __restrict__ out) {
  const size_t total = M * Kp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Kp, c = i % Kp;
    out[i] = c < K ? __ldg(A + r * lda + c) : 0.f;
  }
}

}  // namespace

int gemm_tc(size_t M, size_t N, size_t K, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int ta, int tb,
            int accum, int flags, int passes, cudaStream_t st) {
  // shapes this kernel takes: tall A (not transposed), narrow output, enough rows to fill the machine
  if (ta || N > 256 || N < 1 || K < 1 || M < 4096) return GAI_ERR_UNSUPPORTED;
  if (!encode_fn()) return GAI_ERR_UNSUPPORTED;
  const int n_mma = (int)((N + 15) / 16 * 16);
  const int k_pad = (int)((K + BK - 1) / BK * BK);
  const int num_kb = k_pad / BK;
  const uint32_t b_tile_bytes = (uint32_t)n_mma * BK * 4;
  const uint32_t stage_bytes = 2 * BM * BK * 4 + 2 * b_tile_bytes;  // A_hi | A_lo | B_hi | B_lo   (all multiples of 1024)
  int stages = (int)((200u * 1024u) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return GAI_ERR_UNSUPPORTED;

  // workspace (slot 2): [B_hi | B_lo | padded A (only if A is not TMA-addressable)]
  const bool a_ok = (lda % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
  const size_t kp4 = (K + 3) / 4 * 4;
  const size_t b_elems = (size_t)n_mma * k_pad;
  const size_t ws_bytes = sizeof(float) * (2 * b_elems + (a_ok ? 0 : M * kp4)) + 256;
  void* ws = nullptr;
  int rc = workspace_slot(2, ws_bytes, &ws);
  if (rc != GAI_OK) return rc;
  float* bhi = reinterpret_cast<float*>(ws);
  float* blo = bhi + b_elems;
  prep_b_kernel<<<(unsigned)((b_elems + 255) / 256), 256, 0, st>>>(B, ldb, tb, K, N, k_pad, n_mma, bhi, blo);
  GAI_LAUNCH_CHECK();
  const float* a_src = A;
  size_t a_ld = lda;
  if (!a_ok) {
    float* apad = blo + b_elems + ((64 - ((2 * b_elems) % 64)) % 64);  // keep 256-byte alignment
    size_t blocks = (M * kp4 + 255) / 256;
    const size_t cap = (size_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    pad_a_kernel<<<(unsigned)blocks, 256, 0, st>>>(M, K, kp4, A, lda, apad);
    GAI_LAUNCH_CHECK();
    a_src = apad; a_ld = kp4;
  }
  CUtensorMap map_a, map_bhi, map_blo;
  if (!make_map_f32(&map_a, a_src, M, a_ok ? K : kp4, a_ld, BM, true) || !make_map_f32(&map_bhi, bhi, (uint64_t)n_mma, (uint64_t)k_pad, (uint64_t)k_pad, (uint32_t)n_mma, false) ||
      !make_map_f32(&map_blo, blo, (uint64_t)n_mma, (uint64_t)k_pad, (uint64_t)k_pad, (uint32_t)n_mma, false))
    return set_error(GAI_ERR_CUDA, "gemm_tc", "cuTensorMapEncodeTiled failed");

  TcArgs g;
  g.C = C; g.M = M; g.N = N; g.ldc = ldc; g.n_mma = n_mma; g.num_kb = num_kb; g.stages = stages; g.passes = passes;
  g.accum = accum; g.flags = flags; g.stage_bytes = stage_bytes; g.b_tile_bytes = b_tile_bytes;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static bool configured = false;
  if (!configured) {
    GAI_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    configured = true;
  }
  const size_t tiles = (M + BM - 1) / BM;
  const unsigned grid = (unsigned)(tiles < (size_t)sm_count() ? tiles : (size_t)sm_count());
  gemm_tc_kernel<<<grid, THREADS, smem, st>>>(map_a, map_bhi, map_blo, g);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

}  // namespace gai
