// Internal helpers shared by the sm_100a translation units behind include/gai_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "gai_b200.h"

namespace gai {

int set_error(int code, const char* what, const char* detail);

#define GAI_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) return gai::set_error(GAI_ERR_CUDA, #call, cudaGetErrorString(_e)); \
  } while (0)

#define GAI_CHECK_ARG(cond)                                                        \
  do {                                                                             \
    if (!(cond)) return gai::set_error(GAI_ERR_ARG, "invalid argument", #cond);   \
  } while (0)

extern unsigned long long g_launches;
#define GAI_LAUNCH_CHECK()            \
  do {                                \
    __atomic_fetch_add(&gai::g_launches, 1ull, __ATOMIC_RELAXED); \
    GAI_CUDA(cudaGetLastError());     \
  } while (0)

static inline cudaStream_t S(gai_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (148 on B200); cached.
int sm_count();

// Library-owned scratch (split-K partials, reductions). Grown on demand, never shrunk; one per (device, stream, slot): work on two
// streams (two models, the ranks of an in-process partitioned run, a side stream) never shares a scratch buffer, and growing one
// waits for that stream only.
int workspace(size_t bytes, void** out, cudaStream_t st);          // slot 0: split-K partials, GAT per-vertex scratch, reductions
int workspace_slot(int slot, size_t bytes, void** out, cudaStream_t st);  // slot 1: SpMM padded-input staging, slot 2: dense-transform operands

// Hub threshold: rows longer than a warp's fair share of the edges (nnz / resident warp slots), clamped to [1024, 8192],
// go to the CTA-per-row kernels; everything else is one lane-group per row.
uint32_t hub_degree_for(uint64_t nnz);

// Concatenated dense transforms (gemm_tc.cu): C_j[M x N_j] (+)= sum_p A_p[M x K_p] · op(B_pj), p < nk (K-concatenation: two tall
// operands, one accumulator), j < nn (N-concatenation: one pass over A, two outputs). B_pj is stored [K_p x N_j], or [N_j x K_p] if tb.
struct GemmCat {
  size_t M = 0;
  int nk = 1, nn = 1, tb = 0;
  const float* A[2] = {nullptr, nullptr};
  size_t lda[2] = {0, 0}, K[2] = {0, 0};
  const float* B[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  size_t ldb[2][2] = {{0, 0}, {0, 0}};
  size_t N[2] = {0, 0};
  float* C[2] = {nullptr, nullptr};
  size_t ldc[2] = {0, 0};
  int accum = 0, flags = 0;
  const float* mask = nullptr;  // GAI_EPI_MASK: C = mask > 0 ? C : 0 (nn == 1); with GAI_EPI_BITMASK: uint32 sign-bit words, ldmask in words
  size_t ldmask = 0;
  uint32_t* bits_out = nullptr; // with GAI_EPI_RELU: also write the sign bits of C (one word per row and 32-column chunk)
  size_t ld_bits = 0;
};
// Concatenated weight gradients (gemm_tc_wgrad.cu): C_i = A_i^T · B_i over nrows rows; dual = 1: two A parts (<= 128 columns
// each) against B_0; dual = 2: A_0 against two B parts; dual = 0: the plain product.
struct WgradCat {
  size_t nrows = 0;
  int dual = 0, accum = 0;
  const float* A[2] = {nullptr, nullptr};
  size_t lda[2] = {0, 0}, Kx[2] = {0, 0};
  const float* B[2] = {nullptr, nullptr};
  size_t ldb[2] = {0, 0}, My[2] = {0, 0};
  float* C[2] = {nullptr, nullptr};
  size_t ldc[2] = {0, 0};
};

}  // namespace gai

// Degree-ordered work list of the aggregation kernels over one row range [rb, re) (the whole graph, or one registered segment:
// interior / boundary rows of a 1D partition).
struct gai_worklist {
  uint32_t rb = 0, re = 0;
  uint32_t* row_order = nullptr;  // rows of the range, longest first (ties: ascending id); the first n_hub entries are the hub rows
  uint32_t* claim_ptr = nullptr;  // light rows (row_order + n_hub) cut into claims of <= 32 rows / <= 2048 edges: (begin, end) pairs in execution order
  uint32_t n_hub = 0, n_claims = 0;
};
constexpr int GAI_MAX_SEGMENTS = 8;

// Device graph. Plain struct of device pointers, passed by value into kernels (as the reference passes its
// LearningGraph, include/gnn/graph_operations.h:85).
struct gai_csr {
  uint32_t nv = 0;
  uint64_t nnz = 0;
  uint32_t* rowptr = nullptr;
  uint32_t* colidx = nullptr;
  float* norm_gcn = nullptr;   // 1/sqrt(deg)
  float* norm_mean = nullptr;  // 1/deg
  uint32_t* hub_rows = nullptr;   // == row_order (its first n_hub entries)
  uint32_t n_hub = 0;
  uint32_t hub_degree = 1024;
  uint32_t max_degree = 0;
  uint32_t* row_order = nullptr;  // all rows, longest first (ties: ascending id); the first n_hub entries are the hub rows
  uint32_t* claim_ptr = nullptr;  // light rows (row_order + n_hub) cut into claims of <= 32 rows / <= 2048 edges: (begin, end) pairs in execution order
  uint32_t n_claims = 0;
  unsigned long long* row_counters = nullptr;  // 16 rotating work counters for the persistent light-row kernel
  unsigned counter_seq = 0;
  cudaStream_t aux_stream = nullptr;  // hub-row kernels run here, concurrently with the light-row kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  gai_worklist seg[GAI_MAX_SEGMENTS];  // work lists of registered row segments (gai_csr_set_row_segments)
  int n_seg = 0;
  uint32_t* tperm = nullptr;  // e -> e^T
  bool owns_csr = false;
};
