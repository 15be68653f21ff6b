#include "gai_dist.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "gai_graph.h"

namespace gai_host {

Comm::Comm(int rank, int world, gai_allgather_fn allgather, void* ctx) : rank_(rank), world_(world) {
  die_on(gai_peers_create(rank, world, allgather, ctx, stream(), &peers_), "gai_peers_create");
}
Comm::~Comm() { gai_peers_destroy(peers_); }

int Comm::register_buffer(const void* dptr) {
  int id = -1;
  die_on(gai_peers_register(peers_, const_cast<void*>(dptr), &id), "gai_peers_register");
  ids_[dptr] = id;
  return id;
}
int Comm::id_of(const void* dptr) const {
  auto it = ids_.find(dptr);
  return it == ids_.end() ? -1 : it->second;
}
void Comm::barrier() { die_on(gai_peers_barrier(peers_, stream()), "gai_peers_barrier"); }
void Comm::all_reduce_sum(const float* src, size_t n, float* out) {
  const int id = id_of(src);
  if (id < 0) { std::fprintf(stderr, "Comm::all_reduce_sum: source buffer was never registered\n"); std::exit(EXIT_FAILURE); }
  die_on(gai_peers_combine(peers_, id, n, 1, out, stream()), "gai_peers_combine");
}
void Comm::all_gather(const float* src, size_t n, float* out) {
  const int id = id_of(src);
  if (id < 0) { std::fprintf(stderr, "Comm::all_gather: source buffer was never registered\n"); std::exit(EXIT_FAILURE); }
  die_on(gai_peers_combine(peers_, id, n, 0, out, stream()), "gai_peers_combine");
}
void Comm::check() { die_on(gai_peers_error(peers_, stream()), "gai_peers_error"); }

void ThreadGroup::wait() {
  std::unique_lock<std::mutex> lk(mu);
  const unsigned long long gen = generation;
  if (++arrived == world) {
    arrived = 0;
    generation++;
    cv.notify_all();
  } else {
    cv.wait(lk, [&] { return generation != gen; });
  }
}

void thread_allgather(void* ctx, const void* send, size_t bytes, void* recv_all) {
  ThreadRank* tr = reinterpret_cast<ThreadRank*>(ctx);
  ThreadGroup* g = tr->group;
  g->slots[tr->rank].assign(reinterpret_cast<const unsigned char*>(send), reinterpret_cast<const unsigned char*>(send) + bytes);
  g->wait();
  for (int q = 0; q < g->world; q++) std::memcpy(reinterpret_cast<unsigned char*>(recv_all) + (size_t)q * bytes, g->slots[q].data(), bytes);
  g->wait();  // nobody overwrites its slot before everyone has read it
}

OwnerRange owner_range(uint32_t nv_global, int world, int rank) {
  OwnerRange r;
  r.S = (uint32_t)(((uint64_t)nv_global + world - 1) / world);
  const uint64_t f = (uint64_t)r.S * rank;
  r.first = (uint32_t)(f < nv_global ? f : nv_global);
  r.last = (uint32_t)((uint64_t)r.first + r.S < nv_global ? r.first + r.S : nv_global);
  return r;
}

}  // namespace gai_host
