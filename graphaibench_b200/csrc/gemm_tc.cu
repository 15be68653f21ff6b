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
//   warps 8-11  splitter: rewrites the landed A tile in place as A_hi = A & 0xffffe000 and writes A_lo = (A - A_hi) &
//               0xffffe000 next to it (same swizzled offsets), then fence.proxy.async + arrive.
//   warp 1      MMA issuer: one thread issues 3 x 4 tcgen05.mma (M128 x N x K8) per k-block into one of two TMEM
//               accumulators; tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns per warp, optional "+C" and ReLU, 128-byte row stores.
// The weight is prepared once per call by a tiny kernel (transpose to K-major if needed, zero-pad to [Npad x Kpad],
// split into tf32 hi/lo) so that both operands are K-major and TMA-addressable whatever the caller's layout.
#include <cuda.h>
#include "gai_internal.cuh"

namespace gai {

namespace {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int BK = 32;           // fp32 per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int THREADS = 384;
constexpr uint32_t TF32_MASK = 0xffffe000u;

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused for swizzled K-major, set to 1) in [16,30), stride byte offset = 1024 B (8 rows x 128 B)
// >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

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
            h.x = v.x & TF32_MASK; h.y = v.y & TF32_MASK; h.z = v.z & TF32_MASK; h.w = v.w & TF32_MASK;
            l.x = __float_as_uint(__fsub_rn(__uint_as_float(v.x), __uint_as_float(h.x))) & TF32_MASK;
            l.y = __float_as_uint(__fsub_rn(__uint_as_float(v.y), __uint_as_float(h.y))) & TF32_MASK;
            l.z = __float_as_uint(__fsub_rn(__uint_as_float(v.z), __uint_as_float(h.z))) & TF32_MASK;
            l.w = __float_as_uint(__fsub_rn(__uint_as_float(v.w), __uint_as_float(h.w))) & TF32_MASK;
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
    const uint32_t h = __float_as_uint(v) & TF32_MASK;
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(__float_as_uint(__fsub_rn(v, __uint_as_float(h))) & TF32_MASK);
  }
}

__global__ void pad_a_kernel(size_t M, size_t K, size_t Kp, const float* __restrict__ A, size_t lda, float* __restrict__ out) {
  const size_t total = M * Kp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Kp, c = i % Kp;
    out[i] = c < K ? __ldg(A + r * lda + c) : 0.f;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major [rows x cols] (row stride ld floats), box = [box_rows x 32 floats], SWIZZLE_128B, zero OOB fill.
bool make_map(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, bool stream_once) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, stream_once ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
  if (!make_map(&map_a, a_src, M, a_ok ? K : kp4, a_ld, BM, true) || !make_map(&map_bhi, bhi, (uint64_t)n_mma, (uint64_t)k_pad, (uint64_t)k_pad, (uint32_t)n_mma, false) ||
      !make_map(&map_blo, blo, (uint64_t)n_mma, (uint64_t)k_pad, (uint64_t)k_pad, (uint32_t)n_mma, false))
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
