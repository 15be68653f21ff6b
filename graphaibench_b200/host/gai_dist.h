// 1D-partitioned training, host side: one rank's view of the peer group and of its partition.
//
// The ownership rule is the reference's PartitionedGraph::edgecut_induced_partition1D (src/partitioner/graph_partition.cc:128-178):
// S = ceil(N / P), rank p owns the global ids [p*S, min((p+1)*S, N)); its halo is the set of distinct non-owned neighbours of its masters.
// The reference's induced subgraph numbers masters and halo together in ascending global id (idx_map); here the same two id lists are
// kept — masters first (local row = global id - p*S), then the halo in ascending global id — so that every matrix a rank owns has its
// master rows at the top (dense transforms, loss and optimiser never see the halo) and a halo block below that only aggregations read.
// The data path (csrc/peers.cu) is peer-memory kernels over NVLink: no NCCL call, no pack buffer.
#pragma once
#include <condition_variable>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "gai_b200.h"

namespace gai_host {

class Comm {
 public:
  // `allgather` is the bootstrap collective of gai_peers_create (set-up only); world == 1 needs none.
  Comm(int rank, int world, gai_allgather_fn allgather, void* ctx);
  ~Comm();
  int rank() const { return rank_; }
  int world() const { return world_; }
  bool active() const { return world_ > 1; }
  gai_peers_t peers() const { return peers_; }
  // Collective, same order on every rank. The pointer must be the base of its own device allocation.
  int register_buffer(const void* dptr);
  int id_of(const void* dptr) const;  // -1 if the pointer was never registered
  void barrier();
  // out[i] = sum over ranks of src[i] in rank order (identical bits on every rank); src must be registered, out must not be.
  void all_reduce_sum(const float* src, size_t n, float* out);
  // out[q*n + i] = rank q's src[i]
  void all_gather(const float* src, size_t n, float* out);
  void check();  // exits if a peer never reached a barrier (a rank died)

 private:
  int rank_, world_;
  gai_peers_t peers_ = nullptr;
  std::unordered_map<const void*, int> ids_;
};

// Bootstrap all-gather for ranks that are host threads of one process (gpu_train_* with GAI_PARTS=P): a shared slot table and a
// reusable barrier. One ThreadGroup per run, one (group, rank) context per thread; `thread_allgather` has the gai_allgather_fn signature.
struct ThreadGroup {
  explicit ThreadGroup(int world) : world(world), slots(world) {}
  int world;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long generation = 0;
  std::vector<std::vector<unsigned char>> slots;
  void wait();
};
struct ThreadRank {
  ThreadGroup* group;
  int rank;
};
void thread_allgather(void* thread_rank_ctx, const void* send, size_t bytes, void* recv_all);

// Ownership rule (graph_partition.cc:131-140).
struct OwnerRange {
  uint32_t S, first, last;
};
OwnerRange owner_range(uint32_t nv_global, int world, int rank);

}  // namespace gai_host
