// TEST INFRASTRUCTURE ONLY — C-callable harness around the reference's own partitioner
// (PartitionedGraph::edgecut_induced_partition1D, src/partitioner/graph_partition.cc:128-178), compiled from the
// reference sources by oracle/build_ref.sh into oracle/_ref/libref_part.so. The graph is built in memory with
// GraphT::allocateFrom/fixEndEdge/constructEdge (include/graph.h:142-144); all partition logic is reference code.
#include <cstdint>
#include <cstring>
#include "graph_partition.h"

extern "C" {

struct RefPart {
  Graph* g;
  PartitionedGraph* pg;
};

void* refpart_new(uint32_t nv, const int64_t* rowptr, const uint32_t* colidx, int nparts) {
  RefPart* r = new RefPart();
  r->g = new Graph(nv, (eidType)rowptr[nv]);
  r->g->rowptr()[0] = 0;
  for (uint32_t v = 0; v < nv; v++) r->g->fixEndEdge(v, rowptr[v + 1]);
  for (int64_t e = 0; e < rowptr[nv]; e++) r->g->constructEdge(e, colidx[e]);
  r->pg = new PartitionedGraph(r->g, nparts);
  r->pg->edgecut_induced_partition1D();
  return r;
}
// sizes: [m, ne, local_begin, local_end]
void refpart_sizes(void* h, int part, int64_t* out4) {
  RefPart* r = (RefPart*)h;
  Graph* sg = r->pg->get_subgraph(part);
  out4[0] = sg->V(); out4[1] = sg->E();
  out4[2] = r->pg->get_local_begin(part); out4[3] = r->pg->get_local_end(part);
}
void refpart_get(void* h, int part, uint32_t* idx_map, int64_t* sub_rowptr, uint32_t* sub_colidx) {
  RefPart* r = (RefPart*)h;
  Graph* sg = r->pg->get_subgraph(part);
  memcpy(idx_map, r->pg->idx_map[part].data(), sizeof(uint32_t) * sg->V());
  memcpy(sub_rowptr, sg->rowptr(), sizeof(int64_t) * (sg->V() + 1));
  memcpy(sub_colidx, sg->colidx(), sizeof(uint32_t) * sg->E());
}
}
