// Subgraph sampling for mini-batch ("GraphSAINT-style") training: host-side mirror of the reference's Sampler
// (include/gnn/sampler.h:8-63, src/gnn/sampler.cpp:32-294) — the caller-side step that feeds variable-size graphs into the same layers
// (Model::subgraph_sampling, net.cpp:287-358).
//   select_vertices   frontier sampling over the TRAINING-masked graph (sampler.cpp:170-294): m frontier vertices drawn from the training
//                     set, then n - m steps that pick a frontier vertex with probability proportional to its (clipped) degree through a
//                     flat "dashboard", replace it by a random neighbour and add that neighbour to the set. Sequential and driven by
//                     rand_r: restated operation for operation (the same libc stream gives the same vertex set), on the host.
//   generateSubgraph  mask -> induced subgraph -> re-indexed CSR (sampler.cpp:66-158). Data-parallel: built on the device from the device
//                     CSR of the full graph (csrc/convert.cu: gai_induced_subgraph — id map, per-row count, scan, ordered fill).
#pragma once
#include <set>
#include <vector>
#include "gai_graph.h"

#define DEFAULT_SIZE_FRONTIER 3000  // include/gnn/global.h:31
#define ETA 1.5                     // length factor of the dashboard
#define SAMPLE_CLIP 3000            // degree clip

typedef std::set<index_t> VertexSet;

class Sampler {
 public:
  // g: the full graph (device-resident: copy_to_gpu() must have run), tg: the training-masked graph (host arrays are what the walk reads)
  Sampler(Graph* g, Graph* tg, mask_t* masks, size_t count);
  size_t select_vertices(index_t n, VertexSet& vertex_set, unsigned seed);
  void generateSubgraph(VertexSet& vertex_set, mask_t* masks, Graph* sg);

 protected:
  index_t m;
  size_t count_;
  int avg_deg, subg_deg;
  Graph* full_graph;
  Graph* masked_graph;
  std::vector<index_t> trainingNodes;
};
