#include "gai_sampler.h"
#include <algorithm>
#include <cstdlib>

using gai_host::die_on;
using gai_host::stream;

Sampler::Sampler(Graph* g, Graph* tg, mask_t* masks, size_t count) : m(DEFAULT_SIZE_FRONTIER), count_(count), full_graph(g), masked_graph(tg) {
  for (size_t i = 0; i < full_graph->size(); i++)
    if (masks[i] == 1) trainingNodes.push_back((index_t)i);
  avg_deg = (int)(masked_graph->sizeEdges() / masked_graph->size());
  subg_deg = avg_deg > SAMPLE_CLIP ? SAMPLE_CLIP : avg_deg;
}

namespace {
typedef int db_t;
// The dashboard: one run of `clipped degree` consecutive entries per frontier vertex. For entry j of a run that starts at s and ends
// at e: vertex[j] = the frontier vertex (-1 once it has been replaced), back[j] = s - e at j == s (so that e = s - back[s]) and j - s
// elsewhere (the distance back to s), slot[j] = 1 + index of the vertex in the frontier tables.
struct Dashboard {
  std::vector<db_t> vertex, back, slot;
  void reserve(size_t n) { vertex.reserve(n); back.reserve(n); slot.reserve(n); }
  // sampler.cpp:160-169: when the capacity is short, ask for twice the CURRENT capacity (which may still be short: resize then grows
  // the vector by its own policy) — the capacity decides when the compaction below triggers, so the calls are kept as they are
  void fit(size_t size) {
    if (vertex.capacity() < size) { vertex.reserve(vertex.capacity() * 2); back.reserve(back.capacity() * 2); slot.reserve(slot.capacity() * 2); }
    vertex.resize(size); back.resize(size); slot.resize(size);
  }
  void write_run(db_t start, db_t end, db_t v, db_t slot_id) {
    for (db_t j = start; j < end; j++) { vertex[j] = v; back[j] = j == start ? j - end : j - start; slot[j] = slot_id; }
  }
};
inline db_t clip(db_t d) { return d > SAMPLE_CLIP ? SAMPLE_CLIP : d; }
}  // namespace

size_t Sampler::select_vertices(index_t n, VertexSet& st, unsigned seed) {
  if (n < m) m = n;
  unsigned state = seed;
  auto degree_of = [&](db_t v) { return (db_t)(masked_graph->edge_end_host(v) - masked_graph->edge_begin_host(v)); };
  Dashboard db, fresh;
  db.reserve((size_t)(subg_deg * m * ETA));
  // frontier tables, one entry per vertex that has ever entered the frontier: run length, alive flag, run end, vertex id
  std::vector<db_t> len, alive, end, vid, scan;
  len.reserve(n); alive.reserve(n); end.reserve(n); vid.reserve(n); scan.reserve(n);
  len.resize(m); alive.resize(m); end.resize(m); vid.resize(m);
  for (index_t i = 0; i < m; i++) {
    const db_t v = vid[i] = (db_t)trainingNodes[rand_r(&state) % trainingNodes.size()];
    st.insert((index_t)v);
    len[i] = clip(degree_of(v));
    alive[i] = 1;
    end[i] = 0;
  }
  end[0] = len[0];
  for (index_t i = 1; i < m; i++) end[i] = end[i - 1] + len[i];
  db.fit((size_t)end[m - 1]);
  for (index_t i = 0; i < m; i++) db.write_run(i == 0 ? 0 : end[i - 1], end[i], vid[i], (db_t)i + 1);

  for (index_t itr = 0; itr < n - m; itr++) {
    // a uniformly random live dashboard entry = a frontier vertex with probability proportional to its clipped degree
    db_t pick = -1;
    while (pick == -1) {
      const db_t t = (db_t)(rand_r(&state) % db.vertex.size());
      if ((size_t)t < db.vertex.size() && db.vertex[t] != -1) pick = t;
    }
    pick = db.back[pick] < 0 ? pick : pick - db.back[pick];  // start of its run
    const db_t v = db.vertex[pick];
    const db_t deg = degree_of(v);
    db_t next = deg != 0 ? (db_t)(rand_r(&state) % deg) : -1;
    db_t newlen = 0;
    if (next != -1) {
      next = (db_t)masked_graph->getEdgeDstHost(masked_graph->edge_begin_host(v) + next);
      st.insert((index_t)next);
      alive[db.slot[pick] - 1] = 0;
      len[db.slot[pick] - 1] = 0;
      for (db_t i = pick; i < pick - db.back[pick]; i++) db.vertex[i] = -1;  // retire the run of v
      newlen = clip(degree_of(next));
    }
    if (db.vertex.size() + newlen > db.vertex.capacity()) {
      // compaction: rebuild the dashboard from the live runs, then drop the dead entries of the frontier tables
      scan.resize(len.size());
      scan[0] = len[0];
      for (size_t i = 1; i < len.size(); i++) scan[i] = scan[i - 1] + len[i];
      fresh.vertex.resize(scan.back()); fresh.back.resize(scan.back()); fresh.slot.resize(scan.back());
      end.assign(scan.begin(), scan.end());
      for (size_t i = 0; i < len.size(); i++) {
        if (alive[i] == 0) continue;
        fresh.write_run(i == 0 ? 0 : scan[i - 1], scan[i], vid[i], (db_t)i + 1);
      }
      scan.resize(alive.size());
      scan[0] = alive[0];
      for (size_t i = 1; i < alive.size(); i++) scan[i] = scan[i - 1] + alive[i];  // new slot number of every live entry
      db.vertex.assign(fresh.vertex.begin(), fresh.vertex.end());
      db.back.assign(fresh.back.begin(), fresh.back.end());
      db.slot.assign(fresh.slot.begin(), fresh.slot.end());
      for (auto it = db.slot.begin(); it < db.slot.end(); it++) *it = scan[*it - 1];
      db_t kept = 0;
      for (size_t i = 0; i < len.size(); i++) {
        if (len[i] != 0) { len[kept] = len[i]; alive[kept] = alive[i]; end[kept] = end[i]; vid[kept] = vid[i]; kept++; }
      }
      len.resize(kept); alive.resize(kept); end.resize(kept); vid.resize(kept);
    }
    db.fit(newlen + db.vertex.size());
    len.push_back(newlen);
    alive.push_back(1);
    end.push_back(end.back() + len.back());
    vid.push_back(next);
    db.write_run(*(end.end() - 2), end.back(), vid.back(), (db_t)vid.size());
  }
  return st.size();
}

void Sampler::generateSubgraph(VertexSet& sampledSet, mask_t* masks, Graph* sg) {
  const size_t nv = full_graph->size();
  std::fill(masks, masks + nv, 0);  // createMasks (sampler.h:27-30)
  std::vector<index_t> keep(sampledSet.begin(), sampledSet.end());  // ascending: the new id of a kept vertex is its rank
  for (index_t v : keep) masks[v] = 1;
  if (!full_graph->device()) full_graph->copy_to_gpu();
  void* d_keep = nullptr;
  die_on(gai_malloc(&d_keep, sizeof(index_t) * (keep.size() ? keep.size() : 1)), "gai_malloc");
  die_on(gai_memcpy_h2d(d_keep, keep.data(), sizeof(index_t) * keep.size(), stream()), "gai_memcpy_h2d");
  uint32_t *d_rp = nullptr, *d_ci = nullptr;
  uint64_t nnz = 0;
  die_on(gai_induced_subgraph(full_graph->device(), (uint32_t)keep.size(), (const uint32_t*)d_keep, stream(), &d_rp, &d_ci, &nnz), "gai_induced_subgraph");
  sg->allocateFrom((index_t)keep.size(), (index_t)nnz);
  die_on(gai_memcpy_d2h(sg->row_start_host_ptr(), d_rp, sizeof(index_t) * (keep.size() + 1), stream()), "gai_memcpy_d2h");
  die_on(gai_memcpy_d2h(sg->edge_dst_host_ptr(), d_ci, sizeof(index_t) * nnz, stream()), "gai_memcpy_d2h");
  die_on(gai_stream_sync(stream()), "gai_stream_sync");
  gai_free(d_keep); gai_free(d_rp); gai_free(d_ci);
}
