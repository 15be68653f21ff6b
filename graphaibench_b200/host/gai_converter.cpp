#include "gai_converter.h"
#include <algorithm>
#include <cassert>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include "gai_graph.h"

using gai_host::die_on;
using gai_host::stream;

Converter::Converter(std::string file_type, std::string file_name, bool is_bipartite) {
  if (file_type == "mtx") {
    read_mtx(file_name, is_bipartite);
  } else if (file_type == "edges") {
    read_edgelist(file_name);
  } else if (file_type == "lg" || file_type == "txt") {
    read_lg(file_name);
  } else {
    std::cerr << "unsupported input type " << file_type << " (mtx, edges, lg)\n";
    std::exit(1);
  }
  pairs2CSR();
}

// device: keys -> radix sort -> unique -> offsets (csrc/convert.cu); the two arrays come back to the host for the writers
void Converter::pairs2CSR() {
  const size_t n = psrc.size();
  void *d_src = nullptr, *d_dst = nullptr;
  die_on(gai_malloc(&d_src, sizeof(uint32_t) * (n ? n : 1)), "gai_malloc");
  die_on(gai_malloc(&d_dst, sizeof(uint32_t) * (n ? n : 1)), "gai_malloc");
  die_on(gai_memcpy_h2d(d_src, psrc.data(), sizeof(uint32_t) * n, stream()), "gai_memcpy_h2d");
  die_on(gai_memcpy_h2d(d_dst, pdst.data(), sizeof(uint32_t) * n, stream()), "gai_memcpy_h2d");
  int64_t* d_rp = nullptr;
  uint32_t* d_ci = nullptr;
  uint64_t nnz = 0;
  die_on(gai_coo_to_csr((uint32_t)nv, n, (const uint32_t*)d_src, (const uint32_t*)d_dst, undirected ? 1 : 0, stream(), &d_rp, &d_ci, &nnz), "gai_coo_to_csr");
  ne = (int64_t)nnz;
  rowptr.assign((size_t)nv + 1, 0);
  colidx.assign((size_t)nnz, 0);
  die_on(gai_memcpy_d2h(rowptr.data(), d_rp, sizeof(int64_t) * rowptr.size(), stream()), "gai_memcpy_d2h");
  die_on(gai_memcpy_d2h(colidx.data(), d_ci, sizeof(uint32_t) * colidx.size(), stream()), "gai_memcpy_d2h");
  die_on(gai_stream_sync(stream()), "gai_stream_sync");
  gai_free(d_src); gai_free(d_dst); gai_free(d_rp); gai_free(d_ci);
  psrc.clear(); psrc.shrink_to_fit(); pdst.clear(); pdst.shrink_to_fit();
  gai_host::out() << "|V| " << nv << " |E| " << ne << "\n";
  gai_host::out() << "maximum degree: " << max_degree() << "\n";
}

void Converter::from_pairs(int64_t num_vertices, const uint32_t* src, const uint32_t* dst, size_t n, bool symmetrize) {
  nv = num_vertices; undirected = symmetrize;
  psrc.assign(src, src + n); pdst.assign(dst, dst + n);
  pairs2CSR();
}

uint32_t Converter::max_degree() const {
  int64_t m = 0;
  for (int64_t v = 0; v < nv; v++) m = std::max(m, rowptr[v + 1] - rowptr[v]);
  return (uint32_t)m;
}

void Converter::read_mtx(std::string infile_name, bool is_bipartite) {  // converter.cc:314-420, same checks and exit codes
  gai_host::out() << "Reading MTX file " << infile_name << "\n";
  std::ifstream infile(infile_name.c_str(), std::ios::in);
  std::string start, object, format, field, symmetry, line;
  infile >> start >> object >> format >> field >> symmetry >> std::ws;
  if (start != "%%MatrixMarket") { std::cout << ".mtx file did not start with %%MatrixMarket" << std::endl; std::exit(-21); }
  if ((object != "matrix") || (format != "coordinate")) { std::cout << "only allow matrix coordinate format for .mtx" << std::endl; std::exit(-22); }
  if (field == "complex") { std::cout << "do not support complex weights for .mtx" << std::endl; std::exit(-23); }
  if (field == "pattern") {
    gai_host::out() << "This graph does not have edge weights\n";
  } else if ((field == "real") || (field == "double") || (field == "integer")) {
    std::cout << "weighted .mtx inputs are not supported by this converter (the reference de-duplicates on (neighbour, weight) pairs)" << std::endl;
    std::exit(-27);
  } else { std::cout << "unrecognized field type for .mtx" << std::endl; std::exit(-24); }
  if (symmetry == "symmetric") {
    undirected = true;
    gai_host::out() << "This is a symmetric/undirected graph" << std::endl;
  } else if ((symmetry == "general") || (symmetry == "skew-symmetric")) {
    gai_host::out() << "This is an unsymmetric/directed graph" << std::endl;
    undirected = false;
  } else { std::cout << "unsupported symmetry type for .mtx" << std::endl; std::exit(-25); }
  while (true) {
    char c = infile.peek();
    if (c == '%') infile.ignore(200, '\n'); else break;
  }
  int64_t m, n, nonzeros;
  infile >> m >> n >> nonzeros >> std::ws;
  gai_host::out() << "m=" << m << " n=" << n << " nnz=" << nonzeros << std::endl;
  if (is_bipartite) {
    nv = m + n;
    gai_host::out() << "Bipartite graph\n";
  } else {
    nv = m;
    if (m != n) { std::cout << "matrix must be square for .mtx unless it is a bipartite graph" << std::endl; std::exit(-26); }
  }
  if (is_bipartite) undirected = true;
  psrc.clear(); pdst.clear();
  psrc.reserve((size_t)nonzeros); pdst.reserve((size_t)nonzeros);
  int64_t lines = 0;
  while (std::getline(infile, line)) {
    std::istringstream edge_stream(line);
    int64_t u = 0, v = 0;
    edge_stream >> u;
    edge_stream >> v;
    assert(u > 0 && v > 0);
    int64_t src = u - 1, dst = v - 1;
    if (is_bipartite) dst += m;
    lines++;
    if (src == dst) continue;  // remove selfloops (the device pass drops them as well)
    psrc.push_back((uint32_t)src); pdst.push_back((uint32_t)dst);
  }
  gai_host::out() << "Complete reading " << lines << " lines/edges\n";
}

static void split_ws(const char* line, std::vector<std::string>& out) {
  std::istringstream is(line);
  std::string tok;
  while (is >> tok) out.push_back(tok);
}

void Converter::read_edgelist(std::string infile_name) {  // converter.cc:237-272: 1-based ids, self-loops dropped, both directions kept
  gai_host::out() << "Reading plain edgelist file " << infile_name << "\n";
  std::ifstream infile(infile_name.c_str());
  char line[1024];
  std::vector<std::string> result;
  int64_t num = 0;
  undirected = true;
  while (infile.getline(line, 1024)) {
    result.clear();
    split_ws(line, result);
    if (result.size() < 2) continue;
    int64_t src = atoll(result[0].c_str()), dst = atoll(result[1].c_str());
    if (src < 1 || dst < 1) { std::cout << "vertex ids start from 1 in an edgelist file: src=" << src << " dst=" << dst << "\n"; std::exit(1); }
    src--; dst--;
    if (src == dst) continue;
    num = std::max(num, std::max(src, dst) + 1);
    psrc.push_back((uint32_t)src); pdst.push_back((uint32_t)dst);
  }
  nv = num;
}

void Converter::read_lg(std::string infile_name) {  // converter.cc:274-312: "v id label" / "e from to label" records, first graph of the file
  gai_host::out() << "Reading TXT/LG file " << infile_name << "\n";
  std::ifstream infile(infile_name.c_str());
  char line[1024];
  std::vector<std::string> result;
  int64_t n_labelled = 0;
  bool seen_t = false;
  undirected = true;
  while (infile.getline(line, 1024)) {
    result.clear();
    split_ws(line, result);
    if (result.empty()) continue;
    if (result[0] == "t") {
      if (seen_t && n_labelled) break;  // the next graph of a multi-graph file
      seen_t = true;
    } else if (result[0] == "v" && result.size() >= 3) {
      n_labelled = std::max<int64_t>(n_labelled, atoll(result[1].c_str()) + 1);
    } else if (result[0] == "e" && result.size() >= 4) {
      psrc.push_back((uint32_t)atoll(result[1].c_str())); pdst.push_back((uint32_t)atoll(result[2].c_str()));
    }
  }
  nv = n_labelled;
}

void Converter::generate_binary_graph(std::string outfilename, bool v, bool e, bool, bool) {  // graph.cc:467-508
  gai_host::out() << "Writing graph to file\n";
  if (v) {
    std::ofstream outfile((outfilename + ".vertex.bin").c_str(), std::ios::binary);
    if (!outfile) { std::cout << "File not available\n"; throw 1; }
    outfile.write(reinterpret_cast<const char*>(rowptr.data()), (nv + 1) * sizeof(int64_t));
  }
  if (e) {
    std::ofstream outfile((outfilename + ".edge.bin").c_str(), std::ios::binary);
    if (!outfile) { std::cout << "File not available\n"; throw 1; }
    outfile.write(reinterpret_cast<const char*>(colidx.data()), ne * sizeof(uint32_t));
  }
}

void Converter::write_meta(std::string outfilename) const {
  std::ofstream f((outfilename + ".meta.txt").c_str());
  f << nv << "\n" << ne << "\n" << 4 << " " << 8 << " " << 1 << " " << 2 << "\n" << max_degree() << "\n" << 0 << "\n" << 0 << "\n" << 0 << "\n";
}
