// Multi-GPU exchange over NVLink peer memory for the 1D-partitioned GNN path. No reference counterpart in the GNN
// path (it is single-GPU); the partition rule it serves is PartitionedGraph::edgecut_induced_partition1D
// (src/partitioner/graph_partition.cc:128-178).
//
// One rank per GPU (one process each under torchrun, or one host thread each inside gpu_train_*). Every buffer another rank must read
// — a gathered activation / gradient matrix, a weight-gradient block, the barrier flags — is registered once: the ranks all-gather
// {process id, raw pointer, cudaIpcMemHandle} records through a caller-supplied bootstrap callback, and open the handles of ranks that
// live in other processes (ranks of the same process use the raw pointer). From then on the data path is three kernels, no library
// collective and no staging buffer:
//   peer_barrier_kernel   a flag barrier in peer memory (release store of a sequence number into every peer's flag row, acquire spin
//                         on the own row). Orders "my buffer is complete" before the peers read it and "the peers are done reading"
//                         before it is overwritten. (Ranks that are threads of ONE process may share a device, where a spinning kernel
//                         and an implicitly synchronising call such as cudaMalloc on another rank's thread could wait for each other:
//                         there the barrier is a stream synchronise plus a host rendezvous through the bootstrap callback instead.)
//   halo_pull_kernel      halo rows <- the owners' rows, read straight from the owners' matrices through the mapped pointers into the
//                         halo block of the local matrix: each halo row crosses NVLink exactly once (peer loads bypass the local L2,
//                         B300_MICROARCH.md, so gathering from peer memory inside the aggregation would fetch a row once per EDGE).
//   peer_reduce_kernel    out[i] = sum over ranks, in rank order, of the ranks' buffers (weight gradients; identical bits on every
//                         rank), or the plain concatenation (loss statistics, combined in double on the host).
#include <unistd.h>
#include <vector>
#include "gai_internal.cuh"

constexpr int GAI_MAX_PEERS = 16;
constexpr int GAI_BARRIER_CHANNELS = 2;  // one sequence of barriers per stream that issues them (0: the rank's main stream, 1: its pull stream)

struct gai_peers {
  int rank = 0, world = 1, device = 0;
  gai_allgather_fn allgather = nullptr;
  void* ctx = nullptr;
  unsigned long long pid = 0;
  std::vector<void*> local;               // local base pointer of buffer id
  std::vector<std::vector<void*>> peer;   // peer[id][q]: buffer id of rank q as seen from this device
  std::vector<void*> opened;              // IPC mappings to close
  unsigned long long* flags = nullptr;    // [channels][GAI_MAX_PEERS] this rank's flag rows (buffer id 0)
  unsigned long long** d_flag_rows = nullptr;  // device array [world]: every rank's flag rows
  unsigned long long seq[GAI_BARRIER_CHANNELS] = {0, 0};
  int* d_err = nullptr;
  bool host_barrier = false;  // every rank lives in this process
};

struct gai_halo_plan {
  uint32_t n_halo = 0, S = 0;
  uint32_t seg[GAI_MAX_PEERS + 1] = {0};  // halo rows owned by rank q: [seg[q], seg[q+1])
  uint32_t* d_src_row = nullptr;          // row of halo entry k inside its owner's matrix (global id - owner * S)
};

namespace {

struct IpcRec {
  unsigned long long pid, raw;
  int device, pad;
  cudaIpcMemHandle_t handle;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Thread q: publish `seq` in rank q's flag row (slot = my rank), then wait until rank q has published it in mine. A rank that never
// arrives (crashed peer) trips the cycle budget instead of hanging the GPU: *err is set and the caller reports it.
// Barriers issued on different streams of one rank may execute in either order, and a flag only ever grows within ONE ordered sequence:
// each issuing stream therefore has its own flag row (`channel`) and sequence counter.
__global__ void peer_barrier_kernel(unsigned long long* const* __restrict__ flag_rows, int rank, int world, int channel, unsigned long long seq, int* err) {
  const int q = threadIdx.x;
  if (q >= world) return;
  __threadfence_system();
  st_release_sys(flag_rows[q] + channel * GAI_MAX_PEERS + rank, seq);
  const unsigned long long* mine = flag_rows[rank] + channel * GAI_MAX_PEERS + q;
  const long long t0 = clock64();
  while (ld_acquire_sys(mine) < seq) {
    if (clock64() - t0 > 40000000000ll) { *err = 1; break; }  // ~20 s at 2 GHz
    __nanosleep(64);
  }
  __threadfence_system();
}

struct PullArgs {
  const float* src[GAI_MAX_PEERS];
  uint32_t seg[GAI_MAX_PEERS + 1];
  const uint32_t* src_row;
  float* dst;       // first halo row of the local matrix
  size_t ld_src, ld_dst;
  uint32_t n_halo;
  int world, F, nch;  // nch = float4 chunks per row (vector path)
};

__device__ __forceinline__ int owner_of(const PullArgs& a, uint32_t k) {
  int q = 0;
#pragma unroll 1
  while (q + 1 < a.world && k >= a.seg[q + 1]) q++;
  return q;
}

// One 16-byte chunk per thread and iteration, four independent peer loads in flight per thread (ld.cv: never served from a stale L1 line).
// The grid is cut into world-1 groups of CTAs, group j serving owner (rank + 1 + j) mod world: at every moment a rank reads from ALL its
// peers and every owner's NVLink egress serves all its readers at once. (Walking the halo block in ascending global id instead made
// every rank read owner 0 first, then owner 1, ...: seven readers queued on one GPU's egress while the other links idled — 200 GB/s per
// rank at N = 8 against 580 GB/s at N = 2.)
template <int UNR>
__global__ void __launch_bounds__(256) halo_pull_kernel(const PullArgs a, int rank) {
  const int ngroups = a.world - 1;
  const int j = (int)(blockIdx.x % (unsigned)ngroups);
  const int q = (rank + 1 + j) % a.world;
  const size_t nb = gridDim.x / (unsigned)ngroups;      // CTAs of this group (the launch rounds the grid to a multiple of ngroups)
  const size_t b = blockIdx.x / (unsigned)ngroups;
  const uint32_t k0 = a.seg[q];
  const uint32_t nch = (uint32_t)a.nch;
  const size_t total = (size_t)(a.seg[q + 1] - k0) * nch;   // < 2^32 * 2^16
  const size_t stride = nb * blockDim.x;
  const float* __restrict__ src = a.src[q];
  size_t i = b * blockDim.x + threadIdx.x;
  for (; i + (UNR - 1) * stride < total; i += UNR * stride) {
    float4 v[UNR];
    size_t o[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const size_t idx = i + u * stride;
      const uint32_t kk = (uint32_t)(idx / nch);
      const uint32_t c = (uint32_t)(idx - (size_t)kk * nch);
      const uint32_t k = k0 + kk;
      v[u] = __ldcv(reinterpret_cast<const float4*>(src + (size_t)__ldg(a.src_row + k) * a.ld_src) + c);
      o[u] = (size_t)k * a.ld_dst + (size_t)c * 4;
    }
#pragma unroll
    for (int u = 0; u < UNR; u++) *reinterpret_cast<float4*>(a.dst + o[u]) = v[u];
  }
  for (; i < total; i += stride) {
    const uint32_t kk = (uint32_t)(i / nch);
    const uint32_t c = (uint32_t)(i - (size_t)kk * nch);
    const uint32_t k = k0 + kk;
    *reinterpret_cast<float4*>(a.dst + (size_t)k * a.ld_dst + (size_t)c * 4) =
        __ldcv(reinterpret_cast<const float4*>(src + (size_t)__ldg(a.src_row + k) * a.ld_src) + c);
  }
}
// any width / pitch (per-vertex scalars: degrees, normalisers)
__global__ void halo_pull_scalar_kernel(const PullArgs a) {
  const size_t total = (size_t)a.n_halo * a.F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t k = (uint32_t)(i / a.F);
    const int c = (int)(i - (size_t)k * a.F);
    const int q = owner_of(a, k);
    a.dst[(size_t)k * a.ld_dst + c] = __ldcv(a.src[q] + (size_t)__ldg(a.src_row + k) * a.ld_src + c);
  }
}

struct ReduceArgs {
  const float* src[GAI_MAX_PEERS];
  int world;
};
// sum = 1: out[i] = ((src_0[i] + src_1[i]) + ...) in rank order (the same bits on every rank); sum = 0: out[q * n + i] = src_q[i]
__global__ void peer_reduce_kernel(const ReduceArgs a, size_t n, int sum, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (sum) {
      float s = __ldcv(a.src[0] + i);
      for (int q = 1; q < a.world; q++) s = __fadd_rn(s, __ldcv(a.src[q] + i));
      out[i] = s;
    } else {
      for (int q = 0; q < a.world; q++) out[(size_t)q * n + i] = __ldcv(a.src[q] + i);
    }
  }
}

int launch_barrier(gai_peers* p, cudaStream_t st, int channel = 0) {
  if (p->world == 1) return GAI_OK;
  if (p->host_barrier) {
    GAI_CUDA(cudaStreamSynchronize(st));
    unsigned char token = 0;
    std::vector<unsigned char> all((size_t)p->world);
    p->allgather(p->ctx, &token, 1, all.data());
    return GAI_OK;
  }
  p->seq[channel]++;
  peer_barrier_kernel<<<1, 32, 0, st>>>(p->d_flag_rows, p->rank, p->world, channel, p->seq[channel], p->d_err);
  GAI_LAUNCH_CHECK();
  return GAI_OK;
}

}  // namespace

extern "C" {

int gai_peers_register(gai_peers_t p, void* dptr, int* id_out) {
  GAI_CHECK_ARG(p != nullptr && dptr != nullptr && id_out != nullptr);
  IpcRec mine;
  memset(&mine, 0, sizeof(mine));
  mine.pid = p->pid; mine.raw = (unsigned long long)(uintptr_t)dptr; mine.device = p->device;
  if (p->world > 1) GAI_CUDA(cudaIpcGetMemHandle(&mine.handle, dptr));
  std::vector<IpcRec> all((size_t)p->world);
  if (p->world > 1) p->allgather(p->ctx, &mine, sizeof(IpcRec), all.data());
  else all[0] = mine;
  std::vector<void*> ptrs((size_t)p->world, nullptr);
  for (int q = 0; q < p->world; q++) {
    if (q == p->rank) { ptrs[q] = dptr; continue; }
    if (all[q].pid == p->pid) {
      // a rank of this process (host threads): its pointer is valid here once peer access is on (or it is the same device)
      if (all[q].device != p->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(all[q].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) GAI_CUDA(e);
        cudaGetLastError();
      }
      ptrs[q] = reinterpret_cast<void*>((uintptr_t)all[q].raw);
    } else {
      void* m = nullptr;
      GAI_CUDA(cudaIpcOpenMemHandle(&m, all[q].handle, cudaIpcMemLazyEnablePeerAccess));
      p->opened.push_back(m);
      ptrs[q] = m;
    }
  }
  if (p->local.empty()) {  // first registration (the flag rows, from gai_peers_create): learn whether all ranks share this process
    p->host_barrier = p->world > 1;
    for (int q = 0; q < p->world; q++) p->host_barrier = p->host_barrier && all[q].pid == p->pid;
  }
  *id_out = (int)p->local.size();
  p->local.push_back(dptr);
  p->peer.push_back(ptrs);
  return GAI_OK;
}

int gai_peers_create(int rank, int world, gai_allgather_fn allgather, void* ctx, gai_stream_t stream, gai_peers_t* out) {
  GAI_CHECK_ARG(out != nullptr && world >= 1 && world <= GAI_MAX_PEERS && rank >= 0 && rank < world && (world == 1 || allgather != nullptr));
  gai_peers* p = new gai_peers();
  p->rank = rank; p->world = world; p->allgather = allgather; p->ctx = ctx; p->pid = (unsigned long long)getpid();
  GAI_CUDA(cudaGetDevice(&p->device));
  cudaStream_t st = gai::S(stream);
  GAI_CUDA(cudaMalloc(&p->flags, sizeof(unsigned long long) * GAI_MAX_PEERS * GAI_BARRIER_CHANNELS));
  GAI_CUDA(cudaMemsetAsync(p->flags, 0, sizeof(unsigned long long) * GAI_MAX_PEERS * GAI_BARRIER_CHANNELS, st));
  GAI_CUDA(cudaMalloc(&p->d_err, sizeof(int)));
  GAI_CUDA(cudaMemsetAsync(p->d_err, 0, sizeof(int), st));
  GAI_CUDA(cudaStreamSynchronize(st));  // the flag rows are zero before any peer can see them
  int id = -1;
  int rc = gai_peers_register(p, p->flags, &id);
  if (rc != GAI_OK) { delete p; return rc; }
  GAI_CUDA(cudaMalloc(&p->d_flag_rows, sizeof(void*) * GAI_MAX_PEERS));
  GAI_CUDA(cudaMemcpyAsync(p->d_flag_rows, p->peer[0].data(), sizeof(void*) * world, cudaMemcpyHostToDevice, st));
  GAI_CUDA(cudaStreamSynchronize(st));
  *out = p;
  return GAI_OK;
}

int gai_peers_destroy(gai_peers_t p) {
  if (!p) return GAI_OK;
  for (void* m : p->opened) cudaIpcCloseMemHandle(m);
  cudaFree(p->flags); cudaFree(p->d_err); cudaFree(p->d_flag_rows);
  delete p;
  return GAI_OK;
}

int gai_peers_rank(gai_peers_t p) { return p ? p->rank : 0; }
int gai_peers_world(gai_peers_t p) { return p ? p->world : 1; }

int gai_peers_barrier(gai_peers_t p, gai_stream_t stream) {
  GAI_CHECK_ARG(p != nullptr);
  return launch_barrier(p, gai::S(stream));
}
int gai_peers_barrier_on(gai_peers_t p, int channel, gai_stream_t stream) {
  GAI_CHECK_ARG(p != nullptr && channel >= 0 && channel < GAI_BARRIER_CHANNELS);
  return launch_barrier(p, gai::S(stream), channel);
}

int gai_peers_error(gai_peers_t p, gai_stream_t stream) {
  GAI_CHECK_ARG(p != nullptr);
  int e = 0;
  GAI_CUDA(cudaMemcpyAsync(&e, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, gai::S(stream)));
  GAI_CUDA(cudaStreamSynchronize(gai::S(stream)));
  if (e) return gai::set_error(GAI_ERR_CUDA, "gai_peers_barrier", "a peer did not reach the barrier within the cycle budget");
  return GAI_OK;
}

int gai_halo_plan_create(gai_peers_t p, uint32_t nv_global, uint32_t n_halo, const uint32_t* halo_gids_h, gai_stream_t stream, gai_halo_plan_t* out) {
  GAI_CHECK_ARG(p != nullptr && out != nullptr && (halo_gids_h != nullptr || n_halo == 0));
  gai_halo_plan* h = new gai_halo_plan();
  h->n_halo = n_halo;
  h->S = (uint32_t)(((uint64_t)nv_global + p->world - 1) / p->world);  // graph_partition.cc:131
  std::vector<uint32_t> rows(n_halo ? n_halo : 1);
  int q = 0;
  for (uint32_t k = 0; k < n_halo; k++) {
    const uint32_t g = halo_gids_h[k];
    if (k && g <= halo_gids_h[k - 1]) { delete h; return gai::set_error(GAI_ERR_ARG, "gai_halo_plan_create", "halo ids must be strictly ascending"); }
    const int owner = (int)(g / h->S);
    if (owner >= p->world || owner == p->rank || g >= nv_global) { delete h; return gai::set_error(GAI_ERR_ARG, "gai_halo_plan_create", "halo id owned by this rank or out of range"); }
    while (q < owner) h->seg[++q] = k;
    rows[k] = g - (uint32_t)owner * h->S;
  }
  while (q < GAI_MAX_PEERS) h->seg[++q] = n_halo;
  GAI_CUDA(cudaMalloc(&h->d_src_row, sizeof(uint32_t) * rows.size()));
  GAI_CUDA(cudaMemcpyAsync(h->d_src_row, rows.data(), sizeof(uint32_t) * rows.size(), cudaMemcpyHostToDevice, gai::S(stream)));
  GAI_CUDA(cudaStreamSynchronize(gai::S(stream)));
  *out = h;
  return GAI_OK;
}

int gai_halo_plan_destroy(gai_halo_plan_t h) {
  if (!h) return GAI_OK;
  cudaFree(h->d_src_row);
  delete h;
  return GAI_OK;
}

int gai_halo_pull(gai_peers_t p, gai_halo_plan_t h, int buf_id, int F, size_t ld, float* dst, size_t ld_dst, int flags, gai_stream_t stream) {
  return gai_halo_pull_cols(p, h, buf_id, 0, F, ld, dst, ld_dst, flags, 0, stream);
}

int gai_halo_pull_cols(gai_peers_t p, gai_halo_plan_t h, int buf_id, int col0, int F, size_t ld, float* dst, size_t ld_dst, int flags, int channel,
                       gai_stream_t stream) {
  GAI_CHECK_ARG(p != nullptr && h != nullptr && buf_id > 0 && buf_id < (int)p->local.size() && F > 0 && col0 >= 0);
  GAI_CHECK_ARG(ld >= (size_t)col0 + (size_t)F && ld_dst >= (size_t)col0 + (size_t)F && channel >= 0 && channel < GAI_BARRIER_CHANNELS);
  GAI_CHECK_ARG(dst != nullptr || h->n_halo == 0);
  cudaStream_t st = gai::S(stream);
  if (p->world == 1) return GAI_OK;
  int rc = GAI_OK;
  if (!(flags & GAI_PULL_NO_BARRIER_BEFORE)) { rc = launch_barrier(p, st, channel); if (rc != GAI_OK) return rc; }  // every owner's matrix is complete
  if (h->n_halo) {
    PullArgs a;
    memset(&a, 0, sizeof(a));
    for (int q = 0; q < p->world; q++) a.src[q] = reinterpret_cast<const float*>(p->peer[buf_id][q]) + col0;
    for (int q = 0; q <= GAI_MAX_PEERS; q++) a.seg[q] = h->seg[q];
    a.src_row = h->d_src_row;
    a.dst = dst + col0;
    a.ld_src = ld; a.ld_dst = ld_dst; a.n_halo = h->n_halo; a.world = p->world; a.F = F;
    const bool vec = ld % 4 == 0 && ld_dst % 4 == 0 && col0 % 4 == 0 && reinterpret_cast<uintptr_t>(p->local[buf_id]) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(dst) % 16 == 0;
    if (vec) {
      a.nch = (F + 3) / 4;
      const size_t total = (size_t)h->n_halo * a.nch;
      const size_t ngroups = (size_t)p->world - 1;
      // GAI_PULL_SMALL_GRID: the pull shares the SMs with an aggregation kernel (pipelined exchange): about one CTA per SM with eight
      // 16-byte loads in flight per thread (4.8 MB in flight: NVLink's bandwidth-delay product is ~2.5 MB) instead of filling the machine
      const bool small = (flags & GAI_PULL_SMALL_GRID) != 0;
      const int unr = small ? 8 : 4;
      size_t per_group = (total / ngroups + 256 * unr - 1) / (256 * unr);
      const size_t cap = ((size_t)gai::sm_count() * (small ? 1 : 8) + ngroups - 1) / ngroups;
      if (per_group > cap) per_group = cap;
      if (per_group < 1) per_group = 1;
      if (small) halo_pull_kernel<8><<<(unsigned)(per_group * ngroups), 256, 0, st>>>(a, p->rank);
      else halo_pull_kernel<4><<<(unsigned)(per_group * ngroups), 256, 0, st>>>(a, p->rank);
    } else {
      const size_t total = (size_t)h->n_halo * F;
      size_t blocks = (total + 255) / 256;
      const size_t cap = (size_t)gai::sm_count() * 8;
      if (blocks > cap) blocks = cap;
      halo_pull_scalar_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    }
    GAI_LAUNCH_CHECK();
  }
  if (!(flags & GAI_PULL_NO_BARRIER_AFTER)) rc = launch_barrier(p, st, channel);  // the owners may overwrite their matrices again
  return rc;
}

int gai_peers_combine(gai_peers_t p, int buf_id, size_t n, int sum, float* out, gai_stream_t stream) {
  GAI_CHECK_ARG(p != nullptr && buf_id > 0 && buf_id < (int)p->local.size() && out != nullptr);
  cudaStream_t st = gai::S(stream);
  if (n == 0) return GAI_OK;
  int rc = launch_barrier(p, st);
  if (rc != GAI_OK) return rc;
  ReduceArgs a;
  memset(&a, 0, sizeof(a));
  for (int q = 0; q < p->world; q++) a.src[q] = reinterpret_cast<const float*>(p->peer[buf_id][q]);
  a.world = p->world;
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)gai::sm_count() * 4;
  if (blocks > cap) blocks = cap;
  peer_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, n, sum, out);
  GAI_LAUNCH_CHECK();
  return launch_barrier(p, st);
}

}  // extern "C"
