// Host-side mirror of the reference's text -> binary CSR converter (src/converters/converter.h:73-113, converter.cc) over the device CSR
// builder (csrc/convert.cu: gai_coo_to_csr). Same class name, constructor, reader methods and output files:
//   Converter(file_type, file_name, is_bipartite)   file_type = "mtx" (converter.cc:314-420), "edges" (:237-272) or "lg" (:274-312)
//   generate_binary_graph(prefix, v, e, vl, el)     <prefix>.vertex.bin = int64[nv+1], <prefix>.edge.bin = uint32[ne]
//                                                    (GraphT::write_to_file, src/common/graph.cc:467-508)
// The readers parse on the host exactly as the reference does (1-based ids, self-loops dropped, header checks and exit codes of
// read_mtx); where the reference builds one std::set per vertex and copies it out (adjlist2CSR), the pairs go to the device: keys, radix
// sort, unique, offsets. The reference's own edgelist2CSR is an empty stub (converter.cc:103-104) and its main() only serves the "gr"
// split path, so "edges" / "lg" inputs produce a CSR here and nothing there; "mtx" is pinned bit for bit (tests/test_converter.py).
// Weighted .mtx inputs (the reference de-duplicates on the (neighbour, weight) PAIR and writes an .elabel.bin) are refused.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

class Converter {
 public:
  Converter() {}
  Converter(std::string file_type, std::string file_name, bool is_bipartite);
  void read_edgelist(std::string infile_name);
  void read_lg(std::string infile_name);
  void read_mtx(std::string infile_name, bool is_bipartite);
  // COO pairs already in memory (0-based): the construction the readers end in
  void from_pairs(int64_t num_vertices, const uint32_t* src, const uint32_t* dst, size_t n, bool symmetrize);
  void generate_binary_graph(std::string outfilename, bool v = true, bool e = true, bool vl = true, bool el = true);
  // <outfilename>.meta.txt in the layout GraphT::read_meta_info reads (graph.cc:190-209): the reference leaves this file to the user
  void write_meta(std::string outfilename) const;

  int64_t V() const { return nv; }
  int64_t E() const { return ne; }
  const std::vector<int64_t>& row_offsets() const { return rowptr; }
  const std::vector<uint32_t>& column_indices() const { return colidx; }
  uint32_t max_degree() const;

 private:
  int64_t nv = 0, ne = 0;
  bool undirected = true;
  std::vector<uint32_t> psrc, pdst;  // parsed pairs, 0-based
  std::vector<int64_t> rowptr;
  std::vector<uint32_t> colidx;
  void pairs2CSR();
};
