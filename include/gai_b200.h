/* gai_b200.h — C ABI of the B200-native GNN-layer hot path (drop-in boundary).
 *
 * The reference (chenxuhao/GraphAIBench) has no FFI: its CPU and GPU builds are two object sets behind the
 * same C++ declarations, chosen at link time (src/gnn/Makefile:56-79).  This ABI is the thin layer a third
 * object set calls: every entry point below replaces one reference routine (cited as file:line, paths relative
 * to the reference root) and is what graphaibench_b200/host/*.{h,cpp} — the C++ mirror of the reference's
 * LearningGraph / *_Aggregator / *_layer / loss / optimizer classes — binds to.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (GAI_OK == 0); gai_last_error() gives the text.
 *   - pointers are DEVICE pointers unless the name ends in _h (host).  Matrices are dense row-major fp32 with
 *     an explicit leading dimension where one is given (ld == number of columns for the reference's layout).
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it (no device-wide syncs:
 *     the reference's CudaTest() after every launch, include/utils/cutils.h:18-28, is not reproduced).
 *   - sparse fp32 accumulation is sequential in CSR edge order with a rounded multiply then a rounded add, i.e.
 *     bit-identical to the reference CPU path (src/gnn/gconv/gcn_aggregator.cpp:56-73), for every row length.
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns GAI_ERR_CUDA.
 */
#ifndef GAI_B200_H
#define GAI_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  GAI_OK = 0,
  GAI_ERR_CUDA = 1,        /* a CUDA runtime/driver call failed (includes "no device") */
  GAI_ERR_ARG = 2,         /* invalid argument */
  GAI_ERR_NOMEM = 3,
  GAI_ERR_UNSUPPORTED = 4
};

typedef struct gai_csr* gai_csr_t; /* opaque device graph: replaces LearningGraph's d_* members (include/gnn/lgraph.h:38-44) */
typedef void* gai_stream_t;        /* cudaStream_t */

const char* gai_last_error(void);
int gai_version(void);
int gai_device_count(int* n);
int gai_set_device(int dev);

/* ---- device memory / transfers: replace float_malloc_device, copy_float_device, uint8_malloc_device, ...
 *      (include/utils/math_functions.hh:14-174, src/utilities/math_functions.cu:56-111) ------------------- */
int gai_malloc(void** p, size_t bytes);
int gai_free(void* p);
int gai_memset(void* p, int value, size_t bytes, gai_stream_t stream);
int gai_memcpy_h2d(void* dst, const void* src_h, size_t bytes, gai_stream_t stream);
int gai_memcpy_d2h(void* dst_h, const void* src, size_t bytes, gai_stream_t stream);
int gai_memcpy_d2d(void* dst, const void* src, size_t bytes, gai_stream_t stream);
/* Pitched copy in either direction (host or device pointers, cudaMemcpyDefault): `rows` rows of width_bytes. */
int gai_memcpy2d(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes, size_t width_bytes, size_t rows, gai_stream_t stream);
int gai_stream_sync(gai_stream_t stream);
/* A second stream + cross-stream ordering, for callers that overlap the next step's host->device input copy with the current
 * step's kernels (the reference copies synchronously on the legacy default stream, net.cpp:207-227). */
int gai_stream_create(gai_stream_t* stream);
int gai_stream_destroy(gai_stream_t stream);
int gai_stream_wait_event(gai_stream_t stream, void* ev);
int gai_host_alloc_pinned(void** p_h, size_t bytes);
int gai_host_free_pinned(void* p_h);

/* ---- events + launch accounting (measurement; the reference times with gettimeofday around synchronous ops,
 *      include/timer.h:6-32; here device time is taken with CUDA events on the launching stream) ---------- */
int gai_event_create(void** ev);
int gai_event_record(void* ev, gai_stream_t stream);
int gai_event_elapsed_ms(void* start, void* stop, float* ms); /* synchronises on `stop` */
int gai_event_destroy(void* ev);
uint64_t gai_launch_count(void); /* kernels launched by this library since load (all streams) */

/* ---- host-side graph construction (integer work, bit-exact) ---------------------------------------- */
/* LearningGraph::add_selfloop (include/gnn/lgraph.h:185-218). colidx_out_h has nnz+nv entries. */
int gai_add_selfloop_h(uint32_t nv, const uint32_t* rowptr_h, const uint32_t* colidx_h, uint32_t* rowptr_out_h, uint32_t* colidx_out_h);

/* ---- CSR construction on the device (csrc/convert.cu) ----------------------------------------------------------------
 * COO pairs -> sorted, de-duplicated CSR: what the reference's text converter builds on the host with one std::set per vertex
 * (Converter::read_mtx + adjlist2CSR, src/converters/converter.cc:27-60,314-420): self-loops (and ids >= nv) dropped, with
 * `symmetrize` the reverse of every kept pair added, duplicates removed, rows in ascending neighbour order, int64 row offsets
 * (eidType, as GraphT::write_to_file stores them, src/common/graph.cc:467-508). Inputs and outputs are DEVICE arrays; the two outputs
 * are allocated here (release with gai_free). Synchronises `stream` (the edge count comes back to the host). */
int gai_coo_to_csr(uint32_t nv, uint64_t n_pairs, const uint32_t* src_d, const uint32_t* dst_d, int symmetrize, gai_stream_t stream,
                   int64_t** rowptr_out_d, uint32_t** colidx_out_d, uint64_t* nnz_out);
/* LearningGraph::add_selfloop (lgraph.h:185-218) on the device: row i gains the id first_id + i at its sorted place (first_id = 0 for a
 * whole graph, the first master id for one rank's rows). Out arrays: nv + 1 offsets, nnz + nv columns; must not alias the inputs. */
int gai_add_selfloop_d(uint32_t nv, uint32_t first_id, const uint32_t* rowptr_d, const uint32_t* colidx_d, uint32_t* rowptr_out_d,
                       uint32_t* colidx_out_d, gai_stream_t stream);

/* The subgraph induced by an ascending list of kept vertices, re-indexed to 0..n_keep-1 (new id = rank in the list), neighbour order kept:
 * Sampler::generateSubgraph = getMaskedGraph + reindexSubgraph (src/gnn/sampler.cpp:66-158). Outputs are allocated here (gai_free). */
int gai_induced_subgraph(gai_csr_t g, uint32_t n_keep, const uint32_t* keep_ids_d, gai_stream_t stream, uint32_t** rowptr_out_d, uint32_t** colidx_out_d,
                         uint64_t* nnz_out);

/* ---- device CSR: replaces LearningGraph::alloc_on_device/copy_to_gpu/compute_vertex_data/compute_edge_data
 *      (src/gnn/lgraph.cu:51-140) and the role of GraphGPU::init (include/graph_gpu.h:207-243). -------------
 * Uploads rowptr (u32, nv+1) and colidx (u32, nnz), then on the device computes
 *    norm_gcn[v]  = (float)(1.0 / (double)sqrtf((float)deg_v))   (0 if deg_v == 0)   — lgraph.cpp:22-34
 *    norm_mean[v] = (float)(1.0 / (double)(float)deg_v)                               — sage_aggregator.cpp:17,41
 * and the hub-row list used by the CTA-per-row kernel.  Row lengths are taken from rowptr, so call this on the
 * graph the layers will see (i.e. after add_selfloop for GCN/GAT). */
int gai_csr_create(uint32_t nv, uint64_t nnz, const uint32_t* rowptr_h, const uint32_t* colidx_h, gai_stream_t stream, gai_csr_t* out);
/* Same, from arrays already resident in device memory (borrowed, not copied; must outlive the handle). */
int gai_csr_create_device(uint32_t nv, uint64_t nnz, const uint32_t* rowptr_d, const uint32_t* colidx_d, gai_stream_t stream, gai_csr_t* out);
int gai_csr_destroy(gai_csr_t g);
uint32_t gai_csr_nv(gai_csr_t g);
uint64_t gai_csr_nnz(gai_csr_t g);
const uint32_t* gai_csr_rowptr(gai_csr_t g);   /* LearningGraph::row_start_ptr (lgraph.h:173) */
const uint32_t* gai_csr_colidx(gai_csr_t g);   /* LearningGraph::edge_dst_ptr  (lgraph.h:175) */
const float* gai_csr_vertex_norm(gai_csr_t g); /* LearningGraph::vertex_data_ptr (lgraph.h:179) */
const float* gai_csr_mean_norm(gai_csr_t g);   /* 1 / deg (sage_aggregator.cpp:17,41); with gai_csr_vertex_norm: the arrays a partitioned rank exposes to its peers */
/* Override the per-vertex normalisers with values computed elsewhere (1D partition: norms come from GLOBAL degrees). */
int gai_csr_set_norms(gai_csr_t g, const float* norm_gcn_d, const float* norm_mean_d, gai_stream_t stream);
uint32_t gai_csr_num_hub_rows(gai_csr_t g);
uint32_t gai_csr_max_degree(gai_csr_t g);      /* GraphGPU::get_max_degree (graph_gpu.h:74): longest row */
/* A consumer of the GraphGPU accessor surface (include/gai_graph_gpu.cuh) on this device CSR: the reference's vertex-parallel triangle
 * kernel (src/triangle/gpu_kernels/bs_warp_vertex.cuh) over rows [begin, end): *count_h = sum_v sum_{u in N(v)} |N(v) ∩ N(u)|. */
int gai_triangle_count_rows(gai_csr_t g, uint32_t begin, uint32_t end, uint64_t* count_h, gai_stream_t stream);
/* Register up to 8 row segments [bounds_h[2i], bounds_h[2i+1]) (1D partition: interior rows, boundary rows, all masters; they
 * may overlap). Each gets its own
 * degree-ordered, edge-budgeted work list, used by the *_rows entry points when called with exactly those bounds; any other
 * row range is walked in natural order. Replaces the previous registration. */
int gai_csr_set_row_segments(gai_csr_t g, int n_segments, const uint32_t* bounds_h, gai_stream_t stream);
/* e -> e^T permutation on a structurally symmetric pattern (binary search per edge, as
 * symmetric_csr_transpose, src/utilities/math_functions.cpp:46-74); built once, cached in the handle. */
int gai_csr_build_transpose(gai_csr_t g, gai_stream_t stream);
const uint32_t* gai_csr_transpose_perm(gai_csr_t g);

/* ---- neighbour aggregation (SpMM) -------------------------------------------------------------------
 * out[i, 0:F] = epilogue( sum_{e in row i} w_e * in[col_e, 0:F] ), i in [row_begin, row_end).
 * flags: GAI_EPI_ADD  -> add `addend[i, :]` (ld = ld_out) after the sum;  GAI_EPI_RELU -> max(.,0) last. */
enum { GAI_SPMM_SHARE_SMS = 64 /* aggregations only: size the persistent grid so that one 256-thread CTA of another kernel (the halo pull
                                  of the next column block) fits on every SM next to it */ };
enum { GAI_EPI_NONE = 0, GAI_EPI_RELU = 1, GAI_EPI_ADD = 2, GAI_EPI_MASK = 4 /* dense transforms only: see gai_matmul_kcat */,
       GAI_EPI_BITMASK = 16 /* with GAI_EPI_MASK: `mask` points to uint32 sign-bit words, one per row and 32-column chunk (bit c % 32 of
                               word c / 32 = activation[row, c] > 0), ldmask in words — written by the ReLU epilogue of the layer below */,
       GAI_EPI_PADDED = 8 /* dense transforms only: the rows of C (and of the mask) are padded to a multiple of 4 floats (ldc % 4 == 0,
                             ldc >= round_up(y, 4)); the transform may overwrite the padding columns with zeros (all stores 128-bit) */ };
/* GCN_Aggregator::aggregate == d_aggregate (src/gnn/gconv/gcn_aggregator.cpp:23-77): w_e = norm_i * norm_j. */
int gai_spmm_gcn(gai_csr_t g, int F, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream);
/* SAGE_Aggregator::aggregate (transposed=0, w_e = 1/deg_i) / d_aggregate (transposed=1, w_e = 1/deg_j)
 * (src/gnn/gconv/sage_aggregator.cpp:7-54). */
int gai_spmm_mean(gai_csr_t g, int F, const float* in, int ld_in, float* out, int ld_out, int transposed, int flags, const float* addend, gai_stream_t stream);
/* update_all with explicit per-edge values (src/gnn/gconv/gat_aggregator.cpp:26-45; spmm(), math_functions.cpp:206-219).
 * If perm != NULL the value used for edge e is vals[perm[e]] (transposed attention without materialising it). */
int gai_spmm_edge(gai_csr_t g, int F, const float* vals, const uint32_t* perm, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream);
/* Multi-head edge values (GAT extension, no reference counterpart: the reference has one head): vals[e * heads + h] weighs the columns of
 * head h = column / (F / heads). heads must be a power of two, F / heads a multiple of 4 with a power-of-two number of float4 chunks. */
int gai_spmm_edge_heads(gai_csr_t g, int F, int heads, const float* vals, const uint32_t* perm, const float* in, int ld_in, float* out, int ld_out,
                        int flags, const float* addend, gai_stream_t stream);
/* Aggregation whose epilogue also applies the d_relu of the layer below (out = bit ? out : 0) from sign-bit words written by that
 * layer's ReLU epilogue (GAI_EPI_BITMASK layout): the last op of an aggregate-first layer's backward (gcn_layer.cpp:55-58 followed by
 * gcn_layer.cpp:38-40 of the layer below). */
int gai_spmm_gcn_masked(gai_csr_t g, int F, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend,
                        const uint32_t* mask_bits, int ld_bits, gai_stream_t stream);
int gai_spmm_mean_masked(gai_csr_t g, int F, const float* in, int ld_in, float* out, int ld_out, int transposed, int flags, const float* addend,
                         const uint32_t* mask_bits, int ld_bits, gai_stream_t stream);
/* Every aggregation form over a row range, in one entry point (what the layer classes call). mode: 0 GCN, 1 mean, 2 transposed mean,
 * 3 edge values, 4 edge values through perm. 1D partition: the local CSR's column ids run over [masters | halo]; with in_halo != NULL the
 * rows of neighbour ids >= n_split are read from in_halo[(id - n_split) * ld_in] (the buffer gai_halo_pull fills) and `in` holds master
 * rows only, so no matrix carries a halo block of its own. in_halo == NULL: one matrix, n_split ignored. */
int gai_spmm_rows_ex(gai_csr_t g, int mode, uint32_t row_begin, uint32_t row_end, int F, const float* vals, const uint32_t* perm, const float* in,
                     int ld_in, float* out, int ld_out, int flags, const float* addend, const uint32_t* mask_bits, int ld_bits, const float* in_halo,
                     uint32_t n_split, gai_stream_t stream);
/* Row-range variants for the 1D partition (interior rows first, boundary rows after the halo arrives). */
int gai_spmm_gcn_rows(gai_csr_t g, uint32_t row_begin, uint32_t row_end, int F, const float* in, int ld_in, float* out, int ld_out, int flags, const float* addend, gai_stream_t stream);
int gai_spmm_mean_rows(gai_csr_t g, uint32_t row_begin, uint32_t row_end, int F, const float* in, int ld_in, float* out, int ld_out, int transposed, int flags, const float* addend, gai_stream_t stream);

/* ---- GAT attention (src/gnn/gconv/gat_aggregator.cpp:57-200) ----------------------------------------
 * forward:  t_e = <alpha_l, z_i> + <alpha_r, z_j>;  s_e = LeakyReLU_slope(t_e);  p = softmax over row i;
 *           out_i = sum_e p_e z_j.   temp_scores (t) and norm_scores (p) are saved for the backward (nnz each). */
int gai_gat_forward(gai_csr_t g, int F, const float* z, const float* alpha_l, const float* alpha_r, float slope,
                    float* temp_scores, float* norm_scores, float* out, int flags, gai_stream_t stream);
/* backward: dS_e = <g_i, z_j> (SDDMM); softmax-bwd + LeakyReLU-bwd; d_alpha_l/r (deterministic two-stage reduce,
 *           no atomics); dZ_i = sum_e p^T_e g_j  (transpose through the cached permutation).
 *           dz may alias z (the reference writes dZ over out_temp, gat_layer.cpp:33-36): z is fully consumed first. */
int gai_gat_backward(gai_csr_t g, int F, const float* z, const float* grad_in, float slope, const float* temp_scores,
                     const float* norm_scores, float* scores_grad_ws, float* d_alpha_l, float* d_alpha_r, float* dz, gai_stream_t stream);

/* The same two with explicit row pitches (the layer classes store per-vertex buffers with line-aligned rows, any width F). */
int gai_gat_forward_ld(gai_csr_t g, int F, const float* z, size_t ld_z, const float* alpha_l, const float* alpha_r, float slope,
                       float* temp_scores, float* norm_scores, float* out, size_t ld_out, int flags, gai_stream_t stream);
int gai_gat_backward_ld(gai_csr_t g, int F, const float* z, size_t ld_z, const float* grad_in, size_t ld_grad, float slope, const float* temp_scores,
                        const float* norm_scores, float* scores_grad_ws, float* d_alpha_l, float* d_alpha_r, float* dz, size_t ld_dz,
                        gai_stream_t stream);
/* Multi-head attention (EXTENSION: the reference has one head everywhere, gat_layer.cpp:3-42; BASELINE.json configs[2] names 8). The
 * feature row is cut into `heads` blocks of F / heads columns; head h uses alpha_l / alpha_r entries of its block, has its own scores and
 * row softmax and aggregates its own columns. Score arrays are edge-major, nnz * heads floats: x[e * heads + h]. heads == 1 is the
 * reference path (the *_ld entry points above). heads: a power of two <= 32; heads > 1 needs F / heads % 4 == 0 with a power-of-two
 * number (<= 32) of float4 chunks per head, F <= 512 and 16-byte aligned rows sharing one pitch in the backward pass. */
int gai_gat_forward_heads_ld(gai_csr_t g, int F, int heads, const float* z, size_t ld_z, const float* alpha_l, const float* alpha_r, float slope,
                             float* temp_scores, float* norm_scores, float* out, size_t ld_out, int flags, gai_stream_t stream);
int gai_gat_backward_heads_ld(gai_csr_t g, int F, int heads, const float* z, size_t ld_z, const float* grad_in, size_t ld_grad, float slope,
                              const float* temp_scores, const float* norm_scores, float* scores_grad_ws, float* d_alpha_l, float* d_alpha_r,
                              float* dz, size_t ld_dz, gai_stream_t stream);

/* ---- dense transform: matmul(x,y,z,A,B,C,transA,transB,accum) (src/utilities/math_functions.cpp:142-171;
 *      GPU twin cublasSgemm, math_functions.cu:321-343).  C[x×y] = op(A)[x×z] · op(B)[z×y] (+ C if accum).
 *      fp32 in/out; tensor-core path = tcgen05 kind::tf32 with 3xTF32 error compensation (≈fp32 accuracy).
 *      flags: GAI_EPI_RELU applies max(.,0) in the epilogue.  ---------------------------------------------- */
int gai_matmul(size_t x, size_t y, size_t z, const float* A, const float* B, float* C, int transA, int transB, int accum, int flags, gai_stream_t stream);
/* Same with explicit leading dimensions (lda/ldb are those of the STORED matrices). */
int gai_matmul_ld(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc,
                  int transA, int transB, int accum, int flags, gai_stream_t stream);
/* K-concatenated transform: C[x×y] = A1[x×z1]·op(B1) + A2[x×z2]·op(B2) in ONE pass (one accumulator, C written once),
 * op(B) = B[z×y] or, if transB, B[y×z]^T.  Replaces the reference's sgemm pair with beta = 1 on the second call:
 * SAGE forward  ÂX·W_neigh + X·W_self (src/gnn/gconv/sage_layer.cpp:20-23) and SAGE input gradient
 * dY·W_neigh^T + dZ·W_self^T (sage_layer.cpp:44-52).  flags: GAI_EPI_RELU, and GAI_EPI_MASK = zero C where
 * mask[i,j] <= 0 (d_relu by the forward activation, math_functions.cpp:453-463, applied before C is ever written). */
int gai_matmul_kcat(size_t x, size_t y, size_t z1, const float* A1, size_t lda1, const float* B1, size_t ldb1, size_t z2, const float* A2,
                    size_t lda2, const float* B2, size_t ldb2, float* C, size_t ldc, int transB, int flags, const float* mask, size_t ldmask,
                    uint32_t* relu_bits /* NULL, or with GAI_EPI_RELU: sign-bit words of C written in the same pass */, size_t ld_bits,
                    gai_stream_t stream);
/* Input gradient with the previous layer's d_relu folded into the epilogue: C[x×y] = mask > 0 ? A[x×z]·op(B) : 0
 * (gcn_layer.cpp:51-54 followed by the d_relu of the layer below, gcn_layer.cpp:38-40 / math_functions.cpp:453-463). */
int gai_matmul_mask(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int transB,
                    const float* mask, size_t ldmask, int flags /* GAI_EPI_PADDED, GAI_EPI_BITMASK */, gai_stream_t stream);
/* C[x×y] = ReLU(A[x×z]·B[z×y]) and the sign-bit words of C (ld_bits >= ceil(y / 32)) in one pass. */
int gai_matmul_relu_bits(size_t x, size_t y, size_t z, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, int flags,
                         uint32_t* relu_bits, size_t ld_bits, gai_stream_t stream);
/* N-concatenated transform: C1[x×y1] = A[x×z]·B1[z×y1], C2[x×y2] = A·B2[z×y2] with A read once (SAGE transform-first
 * forward: H·W_neigh for the aggregation and H·W_self for the self term, sage_layer.cpp:26-30). */
int gai_matmul_ncat(size_t x, size_t z, const float* A, size_t lda, size_t y1, const float* B1, size_t ldb1, float* C1, size_t ldc1, size_t y2,
                    const float* B2, size_t ldb2, float* C2, size_t ldc2, int flags /* 0 or GAI_EPI_PADDED */, gai_stream_t stream);
/* Concatenated weight gradients over the same z rows (sage_layer.cpp:37-47, two sgemm(transA) calls in the reference):
 *   gai_wgrad_two_a:  C1[x1×y] = A1[z×x1]^T·B,  C2[x2×y] = A2[z×x2]^T·B    (B read once;  x1, x2 <= 128 on the tensor path)
 *   gai_wgrad_two_b:  C1[x×y1] = A[z×x]^T·B1,   C2[x×y2] = A^T·B2          (A read once) */
int gai_wgrad_two_a(size_t z, size_t y, const float* B, size_t ldb, size_t x1, const float* A1, size_t lda1, float* C1, size_t ldc1, size_t x2,
                    const float* A2, size_t lda2, float* C2, size_t ldc2, gai_stream_t stream);
int gai_wgrad_two_b(size_t z, size_t x, const float* A, size_t lda, size_t y1, const float* B1, size_t ldb1, float* C1, size_t ldc1, size_t y2,
                    const float* B2, size_t ldb2, float* C2, size_t ldc2, gai_stream_t stream);
/* out[i, 0:F] = data[i, 0:F] > 0 ? grad[i, 0:F] : 0 with leading dimensions (d_relu on padded row layouts). */
int gai_d_relu_ld(size_t rows, int F, const float* grad, size_t ld_grad, const float* data, size_t ld_data, float* out, size_t ld_out,
                  gai_stream_t stream);
/* Select the dense path: 0 = auto, 1 = fp32 SIMT FFMA, 2 = tcgen05 3xTF32, 3 = tcgen05 1xTF32 (fast, ~1e-3). */
int gai_set_gemm_mode(int mode);
int gai_get_gemm_mode(void);

/* ---- elementwise (src/utilities/math_functions.cpp:442-463; .cu:242-269) ---------------------------- */
int gai_relu(size_t n, const float* in, float* out, gai_stream_t stream);
int gai_d_relu(size_t n, const float* grad, const float* data, float* out, gai_stream_t stream);
int gai_fill(size_t n, float value, float* out, gai_stream_t stream); /* init_const_gpu, math_functions.cu:12-19 */
/* dropout_cpu / d_dropout_cpu (math_functions.cpp:417-440): out = in * (float)mask * scale with mask ~ Bernoulli(1 - rate) per element.
 * The reference's generator is seeded from /dev/urandom (not reproducible); here mask = hash(seed, call, index): pass a new `call` for
 * every redraw (the reference redraws on every training forward). */
int gai_dropout(size_t n, float rate, float scale, uint64_t seed, uint64_t call, const float* in, uint8_t* mask, float* out, gai_stream_t stream);
int gai_d_dropout(size_t n, float scale, const float* in, const uint8_t* mask, float* out, gai_stream_t stream);

/* ---- l2norm_layer (src/layers/l2norm_layer.cpp:19-64; l2norm/d_l2norm math_functions.cu:158-205) ---- */
int gai_l2norm(int n, int dim, const float* in, float* out, gai_stream_t stream);
int gai_d_l2norm(int n, int dim, const float* feat_in, const float* grad_in, float* grad_out, gai_stream_t stream);
/* Same with explicit row pitches (activation buffers of the layer classes are pitched to a multiple of 4 floats). */
int gai_l2norm_ld(int n, int dim, const float* in, size_t ld_in, float* out, size_t ld_out, gai_stream_t stream);
int gai_d_l2norm_ld(int n, int dim, const float* feat_in, size_t ld_feat, const float* grad_in, size_t ld_grad_in, float* grad_out,
                    size_t ld_grad_out, gai_stream_t stream);

/* ---- softmax_loss_layer (src/layers/softmax_loss_layer.cpp:4-55) + masked_accuracy_single (math_functions.cpp:79-92)
 * forward : rows i in [begin,end) with masks[i]==1 (masks NULL = all): probs_i = softmax(logits_i);
 *           losses[i] = -log(probs_i[label_i]) (log(1e-10) if 0).
 * backward: grad_i = (probs_i - onehot)/(end-begin) on the same rows; other rows untouched.
 * reduce  : stats_d[0] = mean loss over masked rows, stats_d[1] = accuracy (argmax of LOGITS == label), stats_d[2] = count. */
int gai_softmax_ce_forward(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                           float* probs, float* losses, gai_stream_t stream);
int gai_softmax_ce_backward(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* probs,
                            float* grad_out, gai_stream_t stream);
/* Same backward with an explicit denominator and gradient leading dimension: a 1D-partitioned rank owns only part of the
 * reference's [begin,end) range but must scale by the GLOBAL range length (softmax_loss_layer.cpp:31); ld_grad >= ncls lets the
 * gradient land in a 16-byte-aligned row layout the aggregation can gather without a staging copy. */
int gai_softmax_ce_backward_scaled(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* probs,
                                   float* grad_out, int ld_grad, uint64_t denom, gai_stream_t stream);
int gai_masked_loss_accuracy(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                             const float* losses, float* stats_d /*3 floats*/, gai_stream_t stream);
/* The three loss entry points with explicit row pitches for logits / probs / grad (denom = the reference's end - begin). */
int gai_softmax_ce_forward_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                              size_t ld_logits, float* probs, size_t ld_probs, float* losses, gai_stream_t stream);
/* forward + the statistics of gai_masked_loss_accuracy over the same rows in one pass over the logits (forward_prop, net.cpp:458-476,
 * calls the two back to back): stats_d = {mean loss, accuracy, count}. */
int gai_softmax_ce_forward_stats_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                                    size_t ld_logits, float* probs, size_t ld_probs, float* losses, float* stats_d /*3 floats*/, gai_stream_t stream);
int gai_softmax_ce_backward_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* probs,
                               size_t ld_probs, float* grad_out, size_t ld_grad, uint64_t denom, gai_stream_t stream);
int gai_masked_loss_accuracy_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels, const float* logits,
                                size_t ld_logits, const float* losses, float* stats_d /*3 floats*/, gai_stream_t stream);

/* ---- sigmoid_loss_layer (multi-label; src/layers/sigmoid_loss_layer.cpp:4-55, sigmoid / sigmoid_cross_entropy
 *      math_functions.cpp:517-521,553-559) and masked_accuracy_multi = micro-F1 at threshold 0.5 (math_functions.cpp:94-97,580-623).
 * labels_multi: [nv x ncls] multi-hot bytes (Reader::bin_read_vlabels(labels, false), reader.cpp:347-412).
 * forward : probs = sigmoid(logits), losses[i] = sum_j sigmoid cross-entropy, rows of [begin,end) with masks[i]==1.
 * backward: grad = (probs - y) / (float)denom  (the reference divides by end - begin).
 * gai_masked_loss_mean: stats_d = {mean of losses over the masked rows, 0, count}.  gai_masked_f1_micro: *f1_d = micro-F1. */
int gai_sigmoid_ce_forward_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels_multi, const float* logits,
                              size_t ld_logits, float* probs, size_t ld_probs, float* losses, gai_stream_t stream);
int gai_sigmoid_ce_backward_ld(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels_multi, const float* probs,
                               size_t ld_probs, float* grad_out, size_t ld_grad, uint64_t denom, gai_stream_t stream);
int gai_masked_loss_mean(size_t begin, size_t end, const uint8_t* masks, const float* losses, float* stats_d /*3 floats*/, gai_stream_t stream);
int gai_masked_f1_micro(int ncls, size_t begin, size_t end, const uint8_t* masks, const uint8_t* labels_multi, const float* preds, size_t ld_preds,
                        float* f1_d, gai_stream_t stream);

/* ---- adam::update (src/utilities/optimizer.cpp:22-35; GPU twin optimizer.cu:5-36).  The caller owns m, v and
 *      the running powers b1_t/b2_t (they advance once per update() call on the reference's optimiser object). */
int gai_adam_update(size_t n, const float* dW, float* W, float* m, float* v, float lr, float b1, float b2, float b1_t, float b2_t,
                    float eps, gai_stream_t stream);

/* ---- 1D vertex partition: PartitionedGraph::edgecut_induced_partition1D + generate_induced_subgraph
 *      (src/partitioner/graph_partition.cc:70-178), host side, integer, bit-exact.
 * Two-call protocol: with idx_map_h == NULL returns the sizes (*m_out = |masters ∪ halo|, *ne_out = induced nnz);
 * then fills idx_map_h[m], sub_rowptr_h[m+1] (int64, as eidType), sub_colidx_h[ne], local_begin/local_end. */
int gai_partition1d_h(uint32_t nv, const int64_t* rowptr_h, const uint32_t* colidx_h, int nparts, int part, uint32_t* idx_map_h,
                      int64_t* sub_rowptr_h, uint32_t* sub_colidx_h, int64_t* m_out, int64_t* ne_out, uint32_t* local_begin, uint32_t* local_end);

/* ---- multi-GPU: halo exchange and reductions over NVLink peer memory (csrc/peers.cu) ---------------------------------
 * No reference counterpart on the GNN path (single-GPU); the partition these serve follows
 * PartitionedGraph::edgecut_induced_partition1D (src/partitioner/graph_partition.cc:128-178): S = ceil(N / P), rank p owns the
 * global ids [p*S, min((p+1)*S, N)), a rank's matrices hold its masters in ascending global id (row = id - p*S) followed by its halo
 * rows in ascending global id.
 * One rank per GPU: one process each (torchrun) or one host thread each (gpu_train_* with GAI_PARTS). Buffers other ranks read are
 * registered collectively; ranks exchange {pid, pointer, cudaIpcMemHandle} through the caller's bootstrap all-gather. */
typedef struct gai_peers* gai_peers_t;
typedef struct gai_halo_plan* gai_halo_plan_t;
/* Bootstrap collective supplied by the caller (torch.distributed, threads, ...): every rank passes `bytes` bytes, every rank receives
 * the nranks blocks in rank order. Only used at set-up time. */
typedef void (*gai_allgather_fn)(void* ctx, const void* send, size_t bytes, void* recv_all);
int gai_peers_create(int rank, int nranks, gai_allgather_fn allgather, void* ctx, gai_stream_t stream, gai_peers_t* out);
int gai_peers_destroy(gai_peers_t p);
int gai_peers_rank(gai_peers_t p);
int gai_peers_world(gai_peers_t p);
/* Collective: every rank registers its instance of the same logical buffer (a whole cudaMalloc allocation), in the same order. */
int gai_peers_register(gai_peers_t p, void* dptr, int* id_out);
/* Flag barrier in peer memory, enqueued on `stream`: everything this rank enqueued before it is visible to every rank's work behind it. */
int gai_peers_barrier(gai_peers_t p, gai_stream_t stream);
/* The same on barrier channel `channel` (0 or 1). Barriers of one channel must be issued in one order on every rank; a rank that issues
 * barriers from two streams (main stream / pull stream of a pipelined halo exchange) gives each stream its own channel. */
int gai_peers_barrier_on(gai_peers_t p, int channel, gai_stream_t stream);
/* Synchronises `stream`; GAI_ERR_CUDA if a barrier gave up waiting for a peer (a rank died). */
int gai_peers_error(gai_peers_t p, gai_stream_t stream);
/* halo_gids_h: this rank's distinct remote neighbours, strictly ascending global ids (grouped by owner as a consequence). */
int gai_halo_plan_create(gai_peers_t p, uint32_t nv_global, uint32_t n_halo, const uint32_t* halo_gids_h, gai_stream_t stream, gai_halo_plan_t* out);
int gai_halo_plan_destroy(gai_halo_plan_t h);
enum { GAI_PULL_NO_BARRIER_BEFORE = 1, GAI_PULL_NO_BARRIER_AFTER = 2,
       GAI_PULL_SMALL_GRID = 4 /* the pull runs next to another kernel: ~1 CTA per SM, deeper per-thread pipelining */ };
/* dst[k, 0:F] (pitch ld_dst, a private buffer of this rank) <- row of halo vertex k in its owner's instance of buffer `buf_id` (pitch ld_src
 * on every rank), read through the mapped peer pointers. barrier - pull - barrier unless `flags` drops one. */
int gai_halo_pull(gai_peers_t p, gai_halo_plan_t h, int buf_id, int F, size_t ld_src, float* dst, size_t ld_dst, int flags, gai_stream_t stream);
/* Columns [col0, col0 + F) of the same rows (dst still names column 0 of halo row 0); its barriers, if any, run on `channel`. Aggregation
 * is independent per feature column, so an exchange cut into column blocks lets block k be aggregated while block k + 1 crosses NVLink
 * (host/gai_graph.cpp halo_exchange_begin). No reference counterpart (the reference's GNN path is single-GPU). */
int gai_halo_pull_cols(gai_peers_t p, gai_halo_plan_t h, int buf_id, int col0, int F, size_t ld_src, float* dst, size_t ld_dst, int flags, int channel,
                       gai_stream_t stream);
/* sum = 1: out[i] = sum over ranks (rank order, identical bits everywhere) of buffer `buf_id`[i], i < n   (weight-gradient all-reduce);
 * sum = 0: out[q*n + i] = rank q's buffer[i]                                                              (all-gather of small vectors).
 * `out` is a private buffer, not the registered one. barrier - combine - barrier. */
int gai_peers_combine(gai_peers_t p, int buf_id, size_t n, int sum, float* out, gai_stream_t stream);
/* dst_rows[k, 0:F] = src[ids[k], 0:F] (local row gather; `src` may be a peer-mapped pointer). */
int gai_gather_rows(size_t n_ids, const uint32_t* ids, int F, const float* src, int ld_src, float* dst, int ld_dst, gai_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GAI_B200_H */
