// Dense transform on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with TMEM accumulators, operands staged by
// TMA, fp32 in / fp32 out with 3xTF32 error compensation (A·B ≈ A_lo·B_hi + A_hi·B_lo + A_hi·B_hi, every term accumulated
// in fp32 inside TMEM), so that results stay within the fp32 tolerance of the reference's cblas/cublas SGEMM
// (src/utilities/math_functions.cpp:142-171, math_functions.cu:321-343).
//
// Shapes served (row-major, the tall operand is the N x K activation / gradient matrix, the weight is tiny):
//   C[M x N] = A[M x K] · B[K x N]          (forward transforms X·W)           transB = 0
//   C[M x N] = A[M x K] · B[N x K]^T        (input gradients G·W^T)            transB = 1
// with N <= 256 (one MMA tile covers the full output width) and any M, and their concatenated forms (GemmCat):
//   K-concatenation   C = A_0·op(B_0) + A_1·op(B_1)   two tall operands streamed by two TMA maps into ONE accumulator
//                     (SAGE: [ÂX | X]·[W_neigh; W_self], sage_layer.cpp:20-23, and dH = dY·W_n^T + dZ·W_s^T, :44-52)
//   N-concatenation   C_0 = A·B_0, C_1 = A·B_1        one pass over A, two outputs (SAGE transform-first: H·[W_n | W_s])
// The reduction-over-rows product X^T·G (weight gradient) and anything else is declined (GAI_ERR_UNSUPPORTED) and
// served by gemm_tc_wgrad.cu / gemm_simt.cu.
//
// Kernel (persistent, one CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: per k-block (32 fp32 = one 128-byte swizzle row) loads the A tile [128 x 32] and the
//               pre-split weight tiles B_hi/B_lo [N x 32] into a multi-stage shared-memory ring (mbarrier expect_tx).
//   warps 8-11  splitter: rewrites the landed A tile in place as A_hi = rn_tf32(A) and writes A_lo = rn_tf32(A - A_hi)
//               next to it (same swizzled offsets), then fence.proxy.async + arrive.
//   warp 1      MMA issuer: one thread issues 3 x 4 tcgen05.mma (M128 x N x K8) per k-block into one of two TMEM
//               accumulators; tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns per warp, transposed through a padded shared-memory tile so
//               that every global access is a full 128-byte row segment (4 rows per instruction); optional "+C",
//               ReLU, and d_ReLU by a mask matrix (grad = mask > 0 ? grad : 0, math_functions.cpp:453-463).
// The weight is prepared once per call by a tiny kernel (transpose to K-major if needed, zero-pad to [Npad x Kpad],
// split into tf32 hi/lo) so that both operands are K-major and TMA-addressable whatever the caller's layout.
#include <cstdlib>
#include "tc_common.cuh"

namespace gai {

namespace {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int BK = 32;           // fp32 per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int THREADS = 384;

using namespace tc;

constexpr int EPI_PITCH = 36;    // floats per staged row: 32 + 4 keeps 128-bit shared accesses conflict-free both ways
constexpr uint32_t EPI_BYTES = 4 * 32 * EPI_PITCH * 4;  // four epilogue warps

struct TcArgs {
  float* C[2];
  size_t ldc[2];
  int N[2];        // columns of each output; output 1 starts at tile column noff1 (a multiple of 32)
  int noff1, nouts;
  const float* mask;
  size_t ldmask;
  const uint32_t* mask_bits;  // GAI_EPI_BITMASK: the mask as sign bits, one word per (row, 32-column chunk)
  size_t ld_mask_bits;
  uint32_t* bits_out;         // ReLU epilogue also writes the sign bits of C
  size_t ld_bits_out;
  size_t M;
  int n_mma;       // tile width rounded up to a multiple of 16 (UMMA N)
  int nkb0;        // k-blocks of the first K part
  int num_kb;      // k-blocks of 32, both parts
  int last_steps[2];  // k-steps (of 8) the last k-block of each part needs
  int stages;
  int passes;      // 3 = 3xTF32, 1 = single TF32 pass
  int accum, flags;
  uint32_t stage_bytes, b_tile_bytes;
  int debug;       // GAI_TC_DEBUG (timing experiments only, results are wrong): 1 = weights loaded once per stage slot, 2 = no global stores,
                   // 4 = no hi/lo split, 8 = no MMA issue
};

// PAIR: the two CTAs of a cluster (one TPC) work on 256 consecutive rows with ONE cta_group::2 MMA per k-step (M = 256): each CTA stages
// its own 128 rows of A (TMA + hi/lo split as before) and only HALF of the weight tile (its N/2 rows of the K-major B_hi / B_lo), the
// tensor cores of both SMs read both halves. A stage shrinks from 96 KB to 64 KB at N = 256 — three stages instead of two — and the
// L2 -> SM traffic per k-block from 80 KB to 48 KB per SM (the re-streamed weights were 4/5 of it). Protocol: both CTAs' splitter warps
// arrive on the LEADER's conv barrier (count 8, remote mbarrier arrive), the leader's elected thread issues the MMAs and commits with
// multicast onto the empty / accumulator-full barriers of both CTAs, each CTA's epilogue drains its own TMEM half and arrives on the
// leader's accumulator-empty barrier (count 8).
template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, const TcArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], conv_bar[8], empty_bar[8], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;       // 0 = leader of the pair
  const size_t num_rb = (g.M + BM - 1) / BM;                  // 128-row blocks
  const size_t num_tiles = PAIR ? (num_rb + 1) / 2 : num_rb;  // work items: a row block, or a pair of them
  const size_t tile0 = PAIR ? blockIdx.x / 2 : blockIdx.x, tile_step = PAIR ? gridDim.x / 2 : gridDim.x;
  constexpr uint32_t A_BYTES = BM * BK * 4;  // 16 KB

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; i++) { mbar_init(&full_bar[i], 1); mbar_init(&conv_bar[i], PAIR ? 8 : 4); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (PAIR) cluster_sync_all();  // both CTAs resident, barriers initialised, before anything touches the peer
  if (warp == 1) {  // TMEM: 512 columns = two fp32 accumulators of up to 256 columns (the same warp of both CTAs for a pair)
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
  // N >> 3 in bits [17,23), M >> 4 in bits [24,29)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.n_mma >> 3) << 17) | ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);
  auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc_flag) {
    if (PAIR) umma_tf32_pair(d, da, db, idesc, acc_flag); else umma_tf32(d, da, db, idesc, acc_flag);
  };
  auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
  // MMAs of one k-block into accumulator `acc` (one thread), then the commits that release the stage / publish the accumulator
  auto issue_kblock = [&](int kb, int s, int acc) {
    const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
    const uint32_t a_hi = smem_u32(smem + (size_t)s * g.stage_bytes);
    const uint32_t a_lo = a_hi + A_BYTES;
    const uint32_t b_hi = a_hi + 2 * A_BYTES;
    const uint32_t b_lo = b_hi + g.b_tile_bytes;
    // the last k-block of a part holds K mod 32 live columns: the all-zero k-steps behind them are not issued
    const int steps = kb == g.nkb0 - 1 ? g.last_steps[0] : (kb == g.num_kb - 1 ? g.last_steps[1] : BK / 8);
    for (int k = 0; k < ((g.debug & 8) ? 0 : steps); k++) {
      const uint32_t koff = k * 32;  // 8 tf32 = 32 bytes along the swizzled 128-byte row
      const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
      if (g.passes == 3) {
        mma(d_tmem, make_desc_k128(a_lo + koff), make_desc_k128(b_hi + koff), first);
        mma(d_tmem, make_desc_k128(a_hi + koff), make_desc_k128(b_lo + koff), 1u);
        mma(d_tmem, make_desc_k128(a_hi + koff), make_desc_k128(b_hi + koff), 1u);
      } else {
        mma(d_tmem, make_desc_k128(a_hi + koff), make_desc_k128(b_hi + koff), first);
      }
    }
    commit(&empty_bar[s]);  // smem stage reusable once these MMAs have read it (both CTAs of a pair)
    if (kb == g.num_kb - 1) commit(&tfull_bar[acc]);  // accumulator complete
  };
  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t it = 0;
      const int b_row = PAIR ? (int)crank * (g.n_mma / 2) : 0;  // this CTA's half of the weight tile
      for (size_t tile = tile0; tile < num_tiles; tile += tile_step) {
        const size_t rb = PAIR ? tile * 2 + crank : tile;       // a row block past the end loads zeros (TMA out-of-bounds fill)
        for (int kb = 0; kb < g.num_kb; kb++, it++) {
          const int s = it % g.stages;
          mbar_wait(&empty_bar[s], ((it / g.stages) & 1) ^ 1);
          uint8_t* st = smem + (size_t)s * g.stage_bytes;
          const bool load_b = !(g.debug & 1) || it < (uint32_t)g.stages;
          mbar_arrive_expect_tx(&full_bar[s], A_BYTES + (load_b ? (g.passes == 3 ? 2 : 1) * g.b_tile_bytes : 0u));
          if (kb < g.nkb0) tma_load_2d(st, &map_a0, kb * BK, (int)(rb * BM), &full_bar[s]);
          else             tma_load_2d(st, &map_a1, (kb - g.nkb0) * BK, (int)(rb * BM), &full_bar[s]);
          if (load_b) {
            tma_load_2d(st + 2 * A_BYTES, &map_bhi, kb * BK, b_row, &full_bar[s]);
            if (g.passes == 3) tma_load_2d(st + 2 * A_BYTES + g.b_tile_bytes, &map_blo, kb * BK, b_row, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0 && crank == 0) {
      uint32_t it = 0, tcount = 0;
      for (size_t tile = tile0; tile < num_tiles; tile += tile_step, tcount++) {
        const int acc = tcount & 1;
        // epilogue(s) have drained this accumulator
        if (PAIR) mbar_wait_cluster(&tempty_bar[acc], ((tcount >> 1) & 1) ^ 1); else mbar_wait(&tempty_bar[acc], ((tcount >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        for (int kb = 0; kb < g.num_kb; kb++, it++) {
          const int s = it % g.stages;
          if (PAIR) mbar_wait_cluster(&conv_bar[s], (it / g.stages) & 1); else mbar_wait(&conv_bar[s], (it / g.stages) & 1);
          tcgen05_fence_after();
          issue_kblock(kb, s, acc);
        }
      }
    }
  } else if (warp >= 8) {
    // ---------------- splitter: A -> (A_hi in place, A_lo) ----------------
    const int t = threadIdx.x - 256;  // 0..127
    uint32_t it = 0;
    for (size_t tile = tile0; tile < num_tiles; tile += tile_step) {
      for (int kb = 0; kb < g.num_kb; kb++, it++) {
        const int s = it % g.stages;
        mbar_wait(&full_bar[s], (it / g.stages) & 1);
        if (g.passes == 3 && !(g.debug & 4)) {
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
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(&conv_bar[s], 0); else mbar_arrive(&conv_bar[s]); }  // the MMA issuer's barrier
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: TMEM -> registers -> padded smem tile -> full-line global accesses ----------------
    const int q = warp - 4;  // TMEM lane quarter == warp index within the warpgroup
    float* stg = reinterpret_cast<float*>(smem + (size_t)g.stages * g.stage_bytes) + q * (32 * EPI_PITCH);
    const int rsub = lane >> 3, csub = (lane & 7) * 4;  // read-back: 8 lanes cover one 128-byte row segment, 4 rows per pass
    const bool relu = (g.flags & GAI_EPI_RELU) != 0;
    const bool mask_ok = g.mask == nullptr || (((g.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.mask) & 15) == 0));
    uint32_t tcount = 0;
    for (size_t tile = tile0; tile < num_tiles; tile += tile_step, tcount++) {
      const int acc = tcount & 1;
      mbar_wait(&tfull_bar[acc], (tcount >> 1) & 1);
      tcgen05_fence_after();
      const size_t rb = PAIR ? tile * 2 + crank : tile;
      const size_t row0 = rb * BM + (size_t)q * 32;
      const bool tile_full = rb * BM + BM <= g.M;
      const int nrows = tile_full ? 32 : (g.M > row0 ? (int)(g.M - row0 < 32 ? g.M - row0 : 32) : 0);  // live rows of this warp's quarter
      for (int c0 = 0; c0 < g.n_mma; c0 += 32) {
        const int j = (g.nouts > 1 && c0 >= g.noff1) ? 1 : 0;
        const int cbase = c0 - (j ? g.noff1 : 0);
        int ncol = (j ? g.N[1] : g.N[0]) - cbase;  // live columns of this 32-column chunk (warp-uniform)
        if (ncol <= 0) continue;
        if (ncol > 32) ncol = 32;
        float* Cj = j ? g.C[1] : g.C[0];
        const size_t ld = j ? g.ldc[1] : g.ldc[0];
        const bool vec_ok = mask_ok && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cj) & 15) == 0);
        // columns written with 128-bit stores: all of them when the chunk is a whole number of float4 groups, or when the caller
        // declared its rows padded to 4 floats (GAI_EPI_PADDED: the tail group then also covers up to 3 padding columns, which
        // receive the accumulator's zero columns)
        int ncol4 = ncol & ~3;
        if ((g.flags & GAI_EPI_PADDED) && (size_t)(cbase + ((ncol + 3) & ~3)) <= ld) ncol4 = (ncol + 3) & ~3;
        const bool lean = vec_ok && ncol4 >= ncol && !(g.accum && g.mask != nullptr);
        uint32_t r[32];
        if (lean) {
          // ---- lean path: one address per lane, 8 row segments of 16 bytes each, nothing but the epilogue ops in the loop ----
          const bool on = csub < ncol4 && !(g.debug & 2);
          float* cp = Cj + (row0 + rsub) * ld + cbase + csub;
          const size_t step = 4 * ld;
          float4 x[8];  // "+C" or mask operands, requested before the accumulator is read back (their latency overlaps the transpose)
          uint32_t wb[8];  // sign-bit words of the lane's 8 rows (GAI_EPI_BITMASK)
          if (g.accum) {
#pragma unroll
            for (int i = 0; i < 8; i++) if (on && i * 4 + rsub < nrows) x[i] = *reinterpret_cast<const float4*>(cp + i * step);
          } else if (g.mask_bits != nullptr) {
            const uint32_t* bp = g.mask_bits + (row0 + rsub) * g.ld_mask_bits + (cbase >> 5);
#pragma unroll
            for (int i = 0; i < 8; i++) if (i * 4 + rsub < nrows) wb[i] = __ldg(bp + i * 4 * g.ld_mask_bits);
          } else if (g.mask != nullptr) {
            const float* mp = g.mask + (row0 + rsub) * g.ldmask + cbase + csub;
#pragma unroll
            for (int i = 0; i < 8; i++) if (on && i * 4 + rsub < nrows) x[i] = __ldg(reinterpret_cast<const float4*>(mp + i * 4 * g.ldmask));
          }
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)c0, r);
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4)
            *reinterpret_cast<uint4*>(stg + lane * EPI_PITCH + jj) = make_uint4(r[jj], r[jj + 1], r[jj + 2], r[jj + 3]);
          __syncwarp();
          const float* sp = stg + rsub * EPI_PITCH + csub;
          if (g.accum) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
              if (on && i * 4 + rsub < nrows) {
                float4 v = *reinterpret_cast<const float4*>(sp + i * 4 * EPI_PITCH);
                v.x += x[i].x; v.y += x[i].y; v.z += x[i].z; v.w += x[i].w;
                if (relu) { v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f; v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f; }
                *reinterpret_cast<float4*>(cp + i * step) = v;
              }
            }
          } else if (g.mask_bits != nullptr) {
            const int sh = 4 * (lane & 7);  // the lane's four columns within the 32-column word
#pragma unroll
            for (int i = 0; i < 8; i++) {
              if (on && i * 4 + rsub < nrows) {
                float4 v = *reinterpret_cast<const float4*>(sp + i * 4 * EPI_PITCH);
                const uint32_t nib = wb[i] >> sh;
                v.x = (nib & 1u) ? v.x : 0.f; v.y = (nib & 2u) ? v.y : 0.f; v.z = (nib & 4u) ? v.z : 0.f; v.w = (nib & 8u) ? v.w : 0.f;
                *reinterpret_cast<float4*>(cp + i * step) = v;
              }
            }
          } else if (g.bits_out != nullptr) {
            // ReLU + the sign bits of the activation, one 32-bit word per (row, 32-column chunk): the layer above masks its input
            // gradient with 1 bit per element instead of re-reading the activation (78 MB instead of 2.5 GB on C2)
            // each lane collects the 4-bit signs of its 8 rows (nibble i = row i*4 + rsub, its own 4 columns); an 8 x 8 nibble
            // transpose among the 8 lanes of a row group (3 shuffles) then leaves lane l with the full 32-column word of row l*4 + rsub
            uint32_t W = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const bool rv = i * 4 + rsub < nrows;
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (on && rv) v = *reinterpret_cast<const float4*>(sp + i * 4 * EPI_PITCH);
              if (relu) { v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f; v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f; }
              W |= ((v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u)) << (4 * i);
              if (on && rv) *reinterpret_cast<float4*>(cp + i * step) = v;
            }
#pragma unroll
            for (int sidx = 0; sidx < 3; sidx++) {
              const int sft = 4 >> sidx;  // 4, 2, 1
              const uint32_t mlo = sft == 4 ? 0x0000FFFFu : (sft == 2 ? 0x00FF00FFu : 0x0F0F0F0Fu);
              const bool hi = (lane & sft) != 0;
              const uint32_t send = hi ? (W & mlo) : ((W & ~mlo) >> (4 * sft));
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, sft);
              W = hi ? ((W & ~mlo) | recv) : ((W & mlo) | (recv << (4 * sft)));
            }
            const int rl = (lane & 7) * 4 + rsub;
            if (rl < nrows && !(g.debug & 2)) g.bits_out[(row0 + rl) * g.ld_bits_out + (cbase >> 5)] = W;
          } else if (g.mask != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
              if (on && i * 4 + rsub < nrows) {
                float4 v = *reinterpret_cast<const float4*>(sp + i * 4 * EPI_PITCH);
                v.x = x[i].x > 0.f ? v.x : 0.f; v.y = x[i].y > 0.f ? v.y : 0.f; v.z = x[i].z > 0.f ? v.z : 0.f; v.w = x[i].w > 0.f ? v.w : 0.f;
                *reinterpret_cast<float4*>(cp + i * step) = v;
              }
            }
          } else if (tile_full) {
            if (on) {
#pragma unroll
              for (int i = 0; i < 8; i++) {
                float4 v = *reinterpret_cast<const float4*>(sp + i * 4 * EPI_PITCH);
                if (relu) { v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f; v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f; }
                *reinterpret_cast<float4*>(cp + i * step) = v;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
              if (on && i * 4 + rsub < nrows) {
                float4 v = *reinterpret_cast<const float4*>(sp + i * 4 * EPI_PITCH);
                if (relu) { v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f; v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f; }
                *reinterpret_cast<float4*>(cp + i * step) = v;
              }
            }
          }
          __syncwarp();  // the staged tile is rewritten by the next chunk
          continue;
        }
        // ---- general path: ragged column tails, unaligned rows, "+C" together with a mask ----
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)c0, r);
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4)
          *reinterpret_cast<uint4*>(stg + lane * EPI_PITCH + jj) = make_uint4(r[jj], r[jj + 1], r[jj + 2], r[jj + 3]);
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
          const int rl = i * 4 + rsub;
          const size_t row = row0 + rl;
          if (row < g.M && csub < ncol && !(g.debug & 2)) {
            const float4 v = *reinterpret_cast<const float4*>(stg + rl * EPI_PITCH + csub);
            float* cp = Cj + row * ld + cbase + csub;
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (csub + k < ncol) {
                float x = e[k];
                if (g.accum) x += cp[k];
                if (g.mask) x = g.mask[row * g.ldmask + cbase + csub + k] > 0.f ? x : 0.f;
                if (relu) x = x > 0.f ? x : 0.f;
                cp[k] = x;
              }
            }
          }
        }
        __syncwarp();  // the staged tile is rewritten by the next chunk
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_cluster(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
    }
  }

  tcgen05_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();  // a pair: neither CTA leaves while the other may still signal it or read its operands
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Weight preparation: Bt[n][k] = op(B_pj)[k][n] laid out as the tile sees it — K parts back to back (each padded to a
// multiple of 32), output 1's columns from tile column noff1 — zero elsewhere, split into tf32 hi / lo (both K-major).
struct PrepArgs {
  const float* B[2][2];
  size_t ldb[2][2];
  size_t K[2], N[2];
  int tb, nk, nn, noff1, kpad0;
};
__global__ void prep_b_kernel(const PrepArgs a, int k_pad, int n_pad, float* __restrict__ hi, float* __restrict__ lo) {
  const size_t total = (size_t)k_pad * n_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t n = i / k_pad, k = i % k_pad;
    const int j = (a.nn > 1 && n >= (size_t)a.noff1) ? 1 : 0;
    const int p = (a.nk > 1 && k >= (size_t)a.kpad0) ? 1 : 0;
    if (j) n -= a.noff1;
    if (p) k -= a.kpad0;
    float v = 0.f;
    if (n < a.N[j] && k < a.K[p]) v = a.tb ? a.B[p][j][n * a.ldb[p][j] + k] : a.B[p][j][k * a.ldb[p][j] + n];
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

int gemm_tc_cat(const GemmCat& q, int passes, cudaStream_t st) {
  // shapes this kernel takes: tall A (not transposed), narrow output, enough rows to fill the machine
  const size_t M = q.M;
  if (q.nk < 1 || q.nk > 2 || q.nn < 1 || q.nn > 2 || M < 4096) return GAI_ERR_UNSUPPORTED;
  if (q.mask && (q.nn != 1)) return GAI_ERR_UNSUPPORTED;
  for (int p = 0; p < q.nk; p++) if (q.K[p] < 1) return GAI_ERR_UNSUPPORTED;
  for (int j = 0; j < q.nn; j++) if (q.N[j] < 1) return GAI_ERR_UNSUPPORTED;
  const int noff1 = q.nn > 1 ? (int)((q.N[0] + 31) / 32 * 32) : 0;
  const size_t n_tile = q.nn > 1 ? (size_t)noff1 + q.N[1] : q.N[0];
  if (n_tile > 256) return GAI_ERR_UNSUPPORTED;
  if (!encode_fn()) return GAI_ERR_UNSUPPORTED;
  // CTA pairs for wide outputs (three 64 KB stages instead of two 96 KB ones, 48 instead of 80 KB of L2 -> SM traffic per k-block).
  // OPT-IN (GAI_TC_PAIR=1): correct on every transform test, but measured SLOWER than the single-CTA kernel on the C2 epoch (K-concatenated
  // forward 1.77 vs 1.65 ms, masked input gradient 1.17 vs 1.13 ms; profiles/r2_gemm_pair_ab.json) — the depth of the operand ring and the
  // weight re-streaming are therefore NOT what holds these transforms at 0.42 of the HBM roofline.
  const char* pair_env = getenv("GAI_TC_PAIR");   // read per call (the tests switch it)
  const bool allow_pair = pair_env != nullptr && atoi(pair_env) != 0;
  const bool pair = allow_pair && n_tile > 128;
  const int n_mma = pair ? (int)((n_tile + 31) / 32 * 32) : (int)((n_tile + 15) / 16 * 16);
  int nkb[2] = {0, 0}, last_steps[2] = {BK / 8, BK / 8};
  for (int p = 0; p < q.nk; p++) {
    nkb[p] = (int)((q.K[p] + BK - 1) / BK);
    last_steps[p] = (int)((q.K[p] - (size_t)(nkb[p] - 1) * BK + 7) / 8);
  }
  if (q.nk == 1) last_steps[1] = last_steps[0];  // num_kb - 1 == nkb0 - 1: both tests name the same block
  const int num_kb = nkb[0] + nkb[1];
  const int k_pad = num_kb * BK;
  const uint32_t b_rows = pair ? (uint32_t)n_mma / 2 : (uint32_t)n_mma;  // rows of the K-major weight tile one CTA stages
  const uint32_t b_tile_bytes = b_rows * BK * 4;
  const uint32_t stage_bytes = 2 * BM * BK * 4 + 2 * b_tile_bytes;  // A_hi | A_lo | B_hi | B_lo   (all multiples of 1024)
  int stages = (int)((204u * 1024u) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return GAI_ERR_UNSUPPORTED;

  // workspace (slot 2): [B_hi | B_lo | padded copies of the A parts TMA cannot address]
  bool a_ok[2] = {true, true};
  size_t kp4[2] = {0, 0}, pad_elems = 0;
  for (int p = 0; p < q.nk; p++) {
    a_ok[p] = (q.lda[p] % 4 == 0) && (reinterpret_cast<uintptr_t>(q.A[p]) % 16 == 0);
    kp4[p] = (q.K[p] + 3) / 4 * 4;
    if (!a_ok[p]) pad_elems += (M * kp4[p] + 63) / 64 * 64;
  }
  const size_t b_elems = ((size_t)n_mma * k_pad + 63) / 64 * 64;
  void* ws = nullptr;
  int rc = workspace_slot(2, sizeof(float) * (2 * b_elems + pad_elems) + 256, &ws, st);
  if (rc != GAI_OK) return rc;
  float* bhi = reinterpret_cast<float*>(ws);
  float* blo = bhi + b_elems;
  PrepArgs pa;
  memset(&pa, 0, sizeof(pa));
  for (int p = 0; p < q.nk; p++) {
    pa.K[p] = q.K[p];
    for (int j = 0; j < q.nn; j++) { pa.B[p][j] = q.B[p][j]; pa.ldb[p][j] = q.ldb[p][j]; }
  }
  for (int j = 0; j < q.nn; j++) pa.N[j] = q.N[j];
  pa.tb = q.tb; pa.nk = q.nk; pa.nn = q.nn; pa.noff1 = noff1; pa.kpad0 = nkb[0] * BK;
  prep_b_kernel<<<(unsigned)(((size_t)n_mma * k_pad + 255) / 256), 256, 0, st>>>(pa, k_pad, n_mma, bhi, blo);
  GAI_LAUNCH_CHECK();
  CUtensorMap map_a[2], map_bhi, map_blo;
  float* apad = blo + b_elems;
  for (int p = 0; p < 2; p++) {
    const int s = p < q.nk ? p : 0;  // an unused second map mirrors the first (never dereferenced)
    const float* a_src = q.A[s];
    size_t a_ld = q.lda[s], a_cols = q.K[s];
    if (p < q.nk && !a_ok[p]) {
      size_t blocks = (M * kp4[p] + 255) / 256;
      const size_t cap = (size_t)sm_count() * 32;
      if (blocks > cap) blocks = cap;
      pad_a_kernel<<<(unsigned)blocks, 256, 0, st>>>(M, q.K[p], kp4[p], q.A[p], q.lda[p], apad);
      GAI_LAUNCH_CHECK();
      a_src = apad; a_ld = kp4[p]; a_cols = kp4[p];
      apad += (M * kp4[p] + 63) / 64 * 64;
    } else if (p >= q.nk) {
      map_a[1] = map_a[0];
      continue;
    }
    if (!make_map_f32(&map_a[p], a_src, M, a_cols, a_ld, BM, true)) return set_error(GAI_ERR_CUDA, "gemm_tc", "cuTensorMapEncodeTiled failed (A)");
  }
  if (!make_map_f32(&map_bhi, bhi, (uint64_t)n_mma, (uint64_t)k_pad, (uint64_t)k_pad, b_rows, false) ||
      !make_map_f32(&map_blo, blo, (uint64_t)n_mma, (uint64_t)k_pad, (uint64_t)k_pad, b_rows, false))
    return set_error(GAI_ERR_CUDA, "gemm_tc", "cuTensorMapEncodeTiled failed (B)");

  TcArgs g;
  memset(&g, 0, sizeof(g));
  for (int j = 0; j < q.nn; j++) { g.C[j] = q.C[j]; g.ldc[j] = q.ldc[j]; g.N[j] = (int)q.N[j]; }
  g.noff1 = noff1; g.nouts = q.nn; g.ldmask = q.ldmask;
  const bool bitmask = (q.flags & GAI_EPI_MASK) && (q.flags & GAI_EPI_BITMASK);
  g.mask = ((q.flags & GAI_EPI_MASK) && !bitmask) ? q.mask : nullptr;
  g.mask_bits = bitmask ? reinterpret_cast<const uint32_t*>(q.mask) : nullptr; g.ld_mask_bits = q.ldmask;
  g.bits_out = q.bits_out; g.ld_bits_out = q.ld_bits;
  if (bitmask || q.bits_out) {
    // sign-bit words are read / written by the 128-bit epilogue path only: one output, 16-byte aligned rows, every chunk whole
    const bool whole = (q.N[0] % 4 == 0) || ((q.flags & GAI_EPI_PADDED) && q.ldc[0] >= (q.N[0] + 3) / 4 * 4);
    if (q.nn != 1 || q.accum || q.ldc[0] % 4 != 0 || reinterpret_cast<uintptr_t>(q.C[0]) % 16 != 0 || !whole) return GAI_ERR_UNSUPPORTED;
    if (q.bits_out && ((q.flags & GAI_EPI_MASK) || q.ld_bits < (q.N[0] + 31) / 32)) return GAI_ERR_UNSUPPORTED;
    if (bitmask && q.ldmask < (q.N[0] + 31) / 32) return GAI_ERR_UNSUPPORTED;
  }
  g.M = M; g.n_mma = n_mma; g.nkb0 = nkb[0]; g.num_kb = num_kb; g.last_steps[0] = last_steps[0]; g.last_steps[1] = last_steps[1];
  static const int debug_knobs = [] {
    const int k = getenv("GAI_TC_DEBUG") ? atoi(getenv("GAI_TC_DEBUG")) : 0;
    if (k) fprintf(stderr, "libgai_b200: GAI_TC_DEBUG=%d — timing experiment, dense-transform RESULTS ARE WRONG (tools/gemm_probe.py only)\n", k);
    return k;
  }();
  g.debug = debug_knobs;
  g.stages = stages; g.passes = passes; g.accum = q.accum; g.flags = q.flags; g.stage_bytes = stage_bytes; g.b_tile_bytes = b_tile_bytes;
  if ((q.flags & GAI_EPI_MASK) && !q.mask) return set_error(GAI_ERR_ARG, "gemm_tc", "GAI_EPI_MASK without a mask matrix");
  const size_t smem = (size_t)stages * stage_bytes + EPI_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    GAI_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    GAI_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    configured = true;
  }
  const size_t tiles = (M + BM - 1) / BM;
  if (pair) {
    // one cluster of two CTAs per TPC; a pair walks 256-row tiles
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(sm_count() / 2 * 2), 1, 1);
    cfg.blockDim = dim3(THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GAI_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true>, map_a[0], map_a[1], map_bhi, map_blo, g));
    GAI_LAUNCH_CHECK();
    return GAI_OK;
  }
  const unsigned grid = (unsigned)(tiles < (size_t)sm_count() ? tiles : (size_t)sm_count());
  gemm_tc_kernel<false><<<grid, THREADS, smem, st>>>(map_a[0], map_a[1], map_bhi, map_blo, g);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

int gemm_tc(size_t M, size_t N, size_t K, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int ta, int tb,
            int accum, int flags, int passes, cudaStream_t st) {
  if (ta || N > 256 || N < 1 || K < 1) return GAI_ERR_UNSUPPORTED;
  GemmCat q;
  q.M = M; q.A[0] = A; q.lda[0] = lda; q.K[0] = K; q.B[0][0] = B; q.ldb[0][0] = ldb; q.tb = tb;
  q.N[0] = N; q.C[0] = C; q.ldc[0] = ldc; q.accum = accum; q.flags = flags;
  return gemm_tc_cat(q, passes, st);
}

}  // namespace gai
