// Weight gradient on the 5th-generation tensor cores: C[Kx x My] = A^T · B with A [n x Kx] and B [n x My] row-major and
// the reduction running over the n (vertex) rows — dW = X^T·G of every layer (src/gnn/gconv/gcn_layer.cpp:50,56,
// sage_layer.cpp:37-47, gat_layer.cpp:38; matmul(..., transA = true), src/utilities/math_functions.cpp:142-171).
//
// Both operands are "MN-major" for the MMA (the reduction index is the slow axis in memory), which tcgen05 kind::tf32
// takes directly from 128-byte-swizzled shared memory, so the activations are streamed exactly once, in their stored layout:
//   * TMA box = [16 rows x 32 columns] (one 128-byte swizzle span per row, 2 KB); a k-block is 16 rows of A (4 boxes per
//     128-column M tile) and of B (ceil(My/32) boxes). Columns past the matrix edge are zero-filled by TMA.
//   * 3xTF32: splitter warps rewrite each landed box as hi = rn_tf32(x) in place and lo = rn_tf32(x - hi) beside it;
//     the MMA warp issues A_lo·B_hi + A_hi·B_lo + A_hi·B_hi (fp32 accumulation in TMEM) for each 8-row k-group.
//   * split over the rows: a persistent grid of one CTA per SM, each CTA reduces a contiguous range of rows into its own
//     TMEM accumulators (1 or 2 M tiles x up to 256 columns = up to all 512 TMEM columns), writes one fp32 partial, and a
//     second kernel adds the partials in CTA order (deterministic, no atomics).
// HBM-bound by design: bytes = 4·n·(Kx + My), tensor work = 3 · 2·n·128·ceil(Kx/128)·My.
#include "tc_common.cuh"

namespace gai {

namespace {

using namespace tc;

constexpr int WG_BK = 16;                 // rows per k-block
constexpr uint32_t WG_BOX = WG_BK * 128;  // bytes per TMA box
constexpr int WG_THREADS = 384;           // warp 0 TMA, warp 1 MMA, warps 4-11 splitters + epilogue
constexpr int WG_SPLIT_WARPS = 8;

struct WgArgs {
  float* partial;  // [grid][Kx][My]
  size_t nrows;
  size_t blocks_per_cta;  // k-blocks per CTA
  int Kx, My;
  int mt;      // 128-column tiles of A (1 or 2)
  int n_mma;   // My rounded up to a multiple of 32
  int nbox_a, nbox_b;
  int stages, passes;
  uint32_t a_bytes, b_bytes, stage_bytes;
};

// MN-major tf32 operand = layout type SWIZZLE_128B_BASE32B (the only one tcgen05 takes for 32-bit MN-major data): atoms
// of [32 fp32 along M/N] x [4 rows along K] (512 B; 32-byte chunks XOR-swizzled by row, what TMA's SWIZZLE_128B_ATOM_32B
// writes), the next 32 columns one box further (leading byte offset), the next 4 rows 512 B further (stride byte offset).
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(WG_BOX >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)1 << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const WgArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], conv_bar[8], empty_bar[8], done_bar;
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t total_kb = (g.nrows + WG_BK - 1) / WG_BK;
  const size_t kb0 = (size_t)blockIdx.x * g.blocks_per_cta;
  const size_t kb1 = kb0 + g.blocks_per_cta < total_kb ? kb0 + g.blocks_per_cta : total_kb;
  const uint32_t nkb = (uint32_t)(kb1 - kb0);  // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; i++) { mbar_init(&full_bar[i], 1); mbar_init(&conv_bar[i], WG_SPLIT_WARPS); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
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
      for (uint32_t it = 0; it < nkb; it++) {
        const int s = it % g.stages;
        mbar_wait(&empty_bar[s], ((it / g.stages) & 1) ^ 1);
        uint8_t* st = smem + (size_t)s * g.stage_bytes;
        const int row = (int)((kb0 + it) * WG_BK);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(g.nbox_a + g.nbox_b) * WG_BOX);
        for (int c = 0; c < g.nbox_a; c++) tma_load_2d(st + (size_t)c * WG_BOX, &map_a, c * 32, row, &full_bar[s]);
        uint8_t* sb = st + 2 * g.a_bytes;
        for (int c = 0; c < g.nbox_b; c++) tma_load_2d(sb + (size_t)c * WG_BOX, &map_b, c * 32, row, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      // D = F32, A = B = TF32, both MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(g.n_mma >> 3) << 17) | ((128u >> 4) << 24);
      for (uint32_t it = 0; it < nkb; it++) {
        const int s = it % g.stages;
        mbar_wait(&conv_bar[s], (it / g.stages) & 1);
        tcgen05_fence_after();
        const uint32_t a_hi = smem_u32(smem + (size_t)s * g.stage_bytes);
        const uint32_t a_lo = a_hi + g.a_bytes;
        const uint32_t b_hi = a_hi + 2 * g.a_bytes;
        const uint32_t b_lo = b_hi + g.b_bytes;
#pragma unroll
        for (int kg = 0; kg < WG_BK / 8; kg++) {
          const uint32_t koff = kg * 1024;  // 8 rows x 128 B
          for (int t = 0; t < g.mt; t++) {
            const uint32_t toff = (uint32_t)t * 4u * WG_BOX + koff;
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            const uint32_t first = (it == 0 && kg == 0) ? 0u : 1u;
            if (g.passes == 3) {
              umma_tf32(d_tmem, make_desc_mn128(a_lo + toff), make_desc_mn128(b_hi + koff), idesc, first);
              umma_tf32(d_tmem, make_desc_mn128(a_hi + toff), make_desc_mn128(b_lo + koff), idesc, 1u);
              umma_tf32(d_tmem, make_desc_mn128(a_hi + toff), make_desc_mn128(b_hi + koff), idesc, 1u);
            } else {
              umma_tf32(d_tmem, make_desc_mn128(a_hi + toff), make_desc_mn128(b_hi + koff), idesc, first);
            }
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&done_bar);
    }
  } else if (warp >= 4) {
    // ---------------- splitters: x -> (rn_tf32(x) in place, rn_tf32(x - hi)) for the A and B boxes ----------------
    const int t = threadIdx.x - 128;  // 0..255
    const uint32_t a_u4 = g.a_bytes / 16, b_u4 = g.b_bytes / 16;
    // zero the A boxes TMA never fills (columns >= 32*nbox_a of the last M tile), once per stage buffer
    if (g.nbox_a < g.mt * 4) {
      for (int s = 0; s < g.stages; s++) {
        uint4* base = reinterpret_cast<uint4*>(smem + (size_t)s * g.stage_bytes);
        const uint32_t z0 = (uint32_t)g.nbox_a * (WG_BOX / 16);
        for (uint32_t i = z0 + t; i < a_u4; i += WG_SPLIT_WARPS * 32) { base[i] = make_uint4(0, 0, 0, 0); base[a_u4 + i] = make_uint4(0, 0, 0, 0); }
      }
      fence_proxy_async();
    }
    for (uint32_t it = 0; it < nkb; it++) {
      const int s = it % g.stages;
      mbar_wait(&full_bar[s], (it / g.stages) & 1);
      if (g.passes == 3) {
        uint4* ahi = reinterpret_cast<uint4*>(smem + (size_t)s * g.stage_bytes);
        uint4* bhi = ahi + 2 * a_u4;
        const uint32_t a_live = (uint32_t)g.nbox_a * (WG_BOX / 16);
        for (uint32_t i = t; i < a_live + b_u4; i += WG_SPLIT_WARPS * 32) {
          uint4* hi = i < a_live ? ahi + i : bhi + (i - a_live);
          uint4* lo = i < a_live ? hi + a_u4 : hi + b_u4;
          const uint4 v = *hi;
          uint4 h, l;
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
          *hi = h;
          *lo = l;
        }
        fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv_bar[s]);
    }
    // ---------------- epilogue: TMEM -> this CTA's partial ----------------
    const int tile = (warp - 4) >> 2;  // warps 4-7: M tile 0, warps 8-11: M tile 1
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    if (tile < g.mt) {
      mbar_wait(&done_bar, 0);
      tcgen05_fence_after();
      const int kx = tile * 128 + q * 32 + lane;
      float* prow = g.partial + ((size_t)blockIdx.x * g.Kx + (kx < g.Kx ? kx : 0)) * g.My;
      for (int c0 = 0; c0 < g.n_mma; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)tile * 256u + (uint32_t)c0, r);
        if (kx < g.Kx) {
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (c0 + j < g.My) prow[c0 + j] = __uint_as_float(r[j]);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// C = (accum ? C : 0) + sum_p partial[p], p ascending (deterministic).
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, int Kx, int My, size_t ldc, int parts, int accum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kx * My) return;
  const int m = i / My, n = i % My;
  float r = accum ? C[(size_t)m * ldc + n] : 0.0f;
  const size_t stride = (size_t)Kx * My;
  for (int p = 0; p < parts; p++) r += partial[(size_t)p * stride + i];
  C[(size_t)m * ldc + n] = r;
}

// [n x F] (ld) -> [n x Fp], zero-filled tail columns (operands whose row pitch is not a multiple of 16 bytes)
__global__ void wgrad_pad_kernel(size_t n, size_t F, size_t Fp, const float* __restrict__ in, size_t ld, float* __restrict__ out) {
  const size_t total = n * Fp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Fp, c = i % Fp;
    out[i] = c < F ? __ldg(in + r * ld + c) : 0.f;
  }
}

inline bool tma_ok(const float* p, size_t ld) { return (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(p) % 16 == 0); }

}  // namespace

// C[Kx x My] (+)= A^T · B,  A [nrows x Kx] (lda), B [nrows x My] (ldb).
int gemm_tc_wgrad(size_t Kx, size_t My, size_t nrows, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int accum,
                  int flags, int passes, cudaStream_t st) {
  if (Kx < 1 || Kx > 256 || My < 1 || My > 256 || nrows < 4096 || flags != 0) return GAI_ERR_UNSUPPORTED;
  if (!encode_fn()) return GAI_ERR_UNSUPPORTED;
  WgArgs g;
  g.Kx = (int)Kx; g.My = (int)My; g.nrows = nrows; g.passes = passes;
  g.mt = Kx > 128 ? 2 : 1;
  g.n_mma = (int)((My + 31) / 32 * 32);
  g.nbox_a = (int)((Kx + 31) / 32);
  g.nbox_b = g.n_mma / 32;
  g.a_bytes = (uint32_t)g.mt * 4u * WG_BOX;
  g.b_bytes = (uint32_t)g.nbox_b * WG_BOX;
  g.stage_bytes = 2 * (g.a_bytes + g.b_bytes);
  g.stages = (int)((200u * 1024u) / g.stage_bytes);
  if (g.stages > 8) g.stages = 8;
  if (g.stages < 2) return GAI_ERR_UNSUPPORTED;

  const size_t total_kb = (nrows + WG_BK - 1) / WG_BK;
  size_t grid = total_kb < (size_t)sm_count() ? total_kb : (size_t)sm_count();
  g.blocks_per_cta = (total_kb + grid - 1) / grid;
  grid = (total_kb + g.blocks_per_cta - 1) / g.blocks_per_cta;  // every CTA owns at least one k-block

  // workspace slot 0: per-CTA partials; slot 2: padded copies of operands TMA cannot address
  void* ws = nullptr;
  int rc = workspace(sizeof(float) * grid * Kx * My, &ws);
  if (rc != GAI_OK) return rc;
  g.partial = reinterpret_cast<float*>(ws);
  const bool a_ok = tma_ok(A, lda), b_ok = tma_ok(B, ldb);
  const size_t kxp = (Kx + 3) / 4 * 4, myp = (My + 3) / 4 * 4;
  if (!a_ok || !b_ok) {
    void* ws2 = nullptr;
    rc = workspace_slot(2, sizeof(float) * nrows * ((a_ok ? 0 : kxp) + (b_ok ? 0 : myp)) + 512, &ws2);
    if (rc != GAI_OK) return rc;
    float* p = reinterpret_cast<float*>(ws2);
    const size_t cap = (size_t)sm_count() * 32;
    if (!a_ok) {
      size_t blocks = (nrows * kxp + 255) / 256;
      wgrad_pad_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(nrows, Kx, kxp, A, lda, p);
      GAI_LAUNCH_CHECK();
      A = p; lda = kxp;
      p += (nrows * kxp + 63) / 64 * 64;
    }
    if (!b_ok) {
      size_t blocks = (nrows * myp + 255) / 256;
      wgrad_pad_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(nrows, My, myp, B, ldb, p);
      GAI_LAUNCH_CHECK();
      B = p; ldb = myp;
    }
  }
  CUtensorMap map_a, map_b;
  if (!make_map_f32(&map_a, A, nrows, a_ok ? Kx : kxp, lda, WG_BK, true, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) || !make_map_f32(&map_b, B, nrows, b_ok ? My : myp, ldb, WG_BK, true, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return set_error(GAI_ERR_CUDA, "gemm_tc_wgrad", "cuTensorMapEncodeTiled failed");

  const size_t smem = (size_t)g.stages * g.stage_bytes + 1024;
  static bool configured = false;
  if (!configured) {
    GAI_CUDA(cudaFuncSetAttribute(gemm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    configured = true;
  }
  gemm_tc_wgrad_kernel<<<(unsigned)grid, WG_THREADS, smem, st>>>(map_a, map_b, g);
  GAI_LAUNCH_CHECK();
  const int n = (int)(Kx * My);
  wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(g.partial, C, (int)Kx, (int)My, ldc, (int)grid, accum);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

}  // namespace gai
