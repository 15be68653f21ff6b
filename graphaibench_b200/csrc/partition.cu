// 1D vertex partition with master + halo vertices (host side, integer, bit-exact against the reference's
// PartitionedGraph::edgecut_induced_partition1D / generate_induced_subgraph, src/partitioner/graph_partition.cc:70-178).
//
// Partition p of P owns the contiguous master range [p*S, min((p+1)*S, nv)), S = ceil(nv / P). Its vertex set is the
// masters plus every neighbour of a master; local ids enumerate that set in ascending global id, so the masters
// occupy one contiguous local range [local_begin, local_end). The sub-CSR is the subgraph induced on the set
// (halo rows keep only their in-set neighbours), with int64 row offsets as the reference's eidType.
#include <vector>
#include "gai_internal.cuh"

extern "C" int gai_partition1d_h(uint32_t nv, const int64_t* rowptr, const uint32_t* colidx, int nparts, int part, uint32_t* idx_map,
                                 int64_t* sub_rowptr, uint32_t* sub_colidx, int64_t* m_out, int64_t* ne_out, uint32_t* local_begin,
                                 uint32_t* local_end) {
  GAI_CHECK_ARG(rowptr != nullptr && nparts > 0 && part >= 0 && part < nparts);
  GAI_CHECK_ARG(colidx != nullptr || rowptr[nv] == 0);
  const uint64_t S = ((uint64_t)nv + nparts - 1) / nparts;
  const uint64_t first = S * (uint64_t)part < nv ? S * (uint64_t)part : nv;
  const uint64_t last = first + S < nv ? first + S : nv;
  std::vector<uint8_t> in_set(nv, 0);
  for (uint64_t v = first; v < last; v++) {
    in_set[v] = 1;
    for (int64_t e = rowptr[v]; e < rowptr[v + 1]; e++) in_set[colidx[e]] = 1;
  }
  // local id = rank of the vertex inside the set (exclusive scan of the membership flags)
  std::vector<uint32_t> local_id(nv);
  uint64_t m = 0;
  for (uint32_t v = 0; v < nv; v++) {
    local_id[v] = (uint32_t)m;
    m += in_set[v];
  }
  int64_t ne = 0;
  const bool fill = idx_map != nullptr;
  for (uint32_t v = 0; v < nv; v++) {
    if (!in_set[v]) continue;
    const uint32_t lv = local_id[v];
    if (fill) { idx_map[lv] = v; if (sub_rowptr) sub_rowptr[lv] = ne; }
    for (int64_t e = rowptr[v]; e < rowptr[v + 1]; e++) {
      const uint32_t u = colidx[e];
      if (in_set[u]) {
        if (fill && sub_colidx) sub_colidx[ne] = local_id[u];
        ne++;
      }
    }
  }
  if (fill && sub_rowptr) sub_rowptr[m] = ne;
  if (m_out) *m_out = (int64_t)m;
  if (ne_out) *ne_out = ne;
  if (local_begin) *local_begin = first < last ? local_id[first] : 0;
  if (local_end) *local_end = first < last ? local_id[last - 1] + 1 : 0;
  return GAI_OK;
}
