#include "gai_graph.h"
#include "gai_dist.h"
#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

namespace gai_host {
static thread_local gai_stream_t g_stream = nullptr;
gai_stream_t stream() { return g_stream; }
void set_stream(gai_stream_t s) { g_stream = s; }
static thread_local bool g_quiet = false;
void set_quiet(bool quiet) { g_quiet = quiet; }
std::ostream& out() {
  struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
  static NullBuf nb;
  static std::ostream null_stream(&nb);
  return g_quiet ? null_stream : std::cout;
}
void die_on(int status, const char* what) {
  if (status == GAI_OK) return;
  std::fprintf(stderr, "%s failed (status %d): %s\n", what, status, gai_last_error());
  std::exit(EXIT_FAILURE);
}

struct Rec { std::string bucket, shape; double bytes, flops; void* e0; void* e1; };
static thread_local bool g_prof = false;
static thread_local std::vector<Rec> g_recs;
void profile_enable(bool on) { g_prof = on; }
bool profile_enabled() { return g_prof; }
OpScope::OpScope(const char* bucket, const std::string& shape, double bytes, double flops) : idx(-1) {
  if (!g_prof) return;
  Rec r{bucket, shape, bytes, flops, nullptr, nullptr};
  die_on(gai_event_create(&r.e0), "gai_event_create");
  die_on(gai_event_create(&r.e1), "gai_event_create");
  die_on(gai_event_record(r.e0, g_stream), "gai_event_record");
  idx = (int)g_recs.size();
  g_recs.push_back(r);
}
OpScope::~OpScope() {
  if (idx >= 0) die_on(gai_event_record(g_recs[idx].e1, g_stream), "gai_event_record");
}
std::string profile_collect_json() {
  struct Agg { int calls = 0; double ms = 0, bytes = 0, flops = 0; };
  std::vector<std::pair<std::string, Agg>> aggs;
  for (auto& r : g_recs) {
    float ms = 0.f;
    die_on(gai_event_elapsed_ms(r.e0, r.e1, &ms), "gai_event_elapsed_ms");
    gai_event_destroy(r.e0); gai_event_destroy(r.e1);
    const std::string key = r.bucket + "|" + r.shape;
    Agg* a = nullptr;
    for (auto& kv : aggs) if (kv.first == key) a = &kv.second;
    if (!a) { aggs.emplace_back(key, Agg()); a = &aggs.back().second; }
    a->calls++; a->ms += ms; a->bytes += r.bytes; a->flops += r.flops;
  }
  g_recs.clear();
  std::string out = "[";
  char buf[512];
  for (size_t i = 0; i < aggs.size(); i++) {
    const std::string& key = aggs[i].first;
    const size_t bar = key.find('|');
    std::snprintf(buf, sizeof(buf), "%s{\"bucket\":\"%s\",\"shape\":\"%s\",\"calls\":%d,\"ms\":%.6f,\"bytes\":%.0f,\"flops\":%.0f}",
                  i ? "," : "", key.substr(0, bar).c_str(), key.substr(bar + 1).c_str(), aggs[i].second.calls, aggs[i].second.ms,
                  aggs[i].second.bytes, aggs[i].second.flops);
    out += buf;
  }
  return out + "]";
}
}  // namespace gai_host
using gai_host::die_on;

void LearningGraph::allocateFrom(index_t nv, index_t ne) {
  num_vertices_ = nv;
  num_edges_ = ne;
  rowptr_.assign((size_t)nv + 1, 0);
  colidx_.assign(ne, 0);
}

void LearningGraph::add_selfloop() {
  std::vector<index_t> rp((size_t)num_vertices_ + 1), ci((size_t)num_edges_ + num_vertices_);
  die_on(gai_add_selfloop_h(num_vertices_, rowptr_.data(), colidx_.data(), rp.data(), ci.data()), "gai_add_selfloop_h");
  rowptr_.swap(rp);
  colidx_.swap(ci);
  num_edges_ += num_vertices_;
}

void LearningGraph::degree_counting() {
  max_degree = 0;
  for (index_t v = 0; v < num_vertices_; v++) max_degree = std::max(max_degree, rowptr_[v + 1] - rowptr_[v]);
}

LearningGraph* LearningGraph::generate_masked_graph(mask_t* masks) {
  // keep vertex ids, keep only edges whose two endpoints are in the mask (lgraph.h:231-272)
  LearningGraph* mg = new LearningGraph(is_device);
  std::vector<index_t> rp((size_t)num_vertices_ + 1, 0);
  for (index_t v = 0; v < num_vertices_; v++) {
    index_t d = 0;
    if (masks[v] == 1)
      for (index_t e = rowptr_[v]; e < rowptr_[v + 1]; e++) d += masks[colidx_[e]] == 1;
    rp[v + 1] = rp[v] + d;
  }
  mg->allocateFrom(num_vertices_, rp[num_vertices_]);
  for (index_t v = 0; v < num_vertices_; v++) {
    mg->fixEndEdge(v, rp[v + 1]);
    if (masks[v] != 1) continue;
    index_t k = rp[v];
    for (index_t e = rowptr_[v]; e < rowptr_[v + 1]; e++)
      if (masks[colidx_[e]] == 1) mg->constructEdge(k++, colidx_[e]);
  }
  gai_host::out() << "masked graph: num_vertices = " << mg->size() << ", num_edges = " << mg->sizeEdges() << "\n";
  return mg;
}

void LearningGraph::alloc_on_device() {}
void LearningGraph::alloc_on_device(index_t) {}

void LearningGraph::copy_to_gpu() {
  if (dev_) { gai_csr_destroy(dev_); dev_ = nullptr; }
  if (!comm_) {
    die_on(gai_csr_create(num_vertices_, num_edges_, rowptr_.data(), colidx_.data(), gai_host::stream(), &dev_), "gai_csr_create");
    return;
  }
  // partitioned: the device CSR spans masters + halo (halo rows are empty); normalisers come from GLOBAL degrees — the masters' rows
  // are complete, so their degrees are global already; the halo vertices' normalisers are read from their owners
  const size_t m = rows_with_halo();
  std::vector<index_t> rp(m + 1);
  std::copy(rowptr_.begin(), rowptr_.end(), rp.begin());
  std::fill(rp.begin() + num_vertices_ + 1, rp.end(), rowptr_[num_vertices_]);
  die_on(gai_csr_create((index_t)m, num_edges_, rp.data(), colidx_.data(), gai_host::stream(), &dev_), "gai_csr_create");
  const index_t seg[2] = {0, num_vertices_};
  die_on(gai_csr_set_row_segments(dev_, 1, seg, gai_host::stream()), "gai_csr_set_row_segments");
  if (plan_) gai_halo_plan_destroy(plan_);
  die_on(gai_halo_plan_create(comm_->peers(), nv_global_, (uint32_t)halo_gids_.size(), halo_gids_.data(), gai_host::stream(), &plan_), "gai_halo_plan_create");
  const float* norms[2] = {gai_csr_vertex_norm(dev_), gai_csr_mean_norm(dev_)};
  for (const float* nrm : norms) {
    const int id = comm_->register_buffer(nrm);
    die_on(gai_halo_pull(comm_->peers(), plan_, id, 1, 1, const_cast<float*>(nrm) + num_vertices_, 1, 0, gai_host::stream()), "gai_halo_pull(norms)");
  }
}

void LearningGraph::add_selfloop_rows(index_t first) {
  // lgraph.h:185-218 on rows [first, first + n): the loop id of row r is first + r
  const index_t n = num_vertices_;
  std::vector<index_t> rp((size_t)n + 1), ci((size_t)num_edges_ + n);
  for (index_t r = 0; r < n; r++) {
    const index_t b = rowptr_[r], e = rowptr_[r + 1], self = first + r;
    index_t pos = e;
    for (index_t k = b; k < e; k++)
      if (colidx_[k] > self) { pos = k; break; }
    index_t* o = ci.data() + (size_t)b + r;
    std::copy(colidx_.begin() + b, colidx_.begin() + pos, o);
    o[pos - b] = self;
    std::copy(colidx_.begin() + pos, colidx_.begin() + e, o + (pos - b) + 1);
    rp[r] = b + r;
  }
  rp[n] = rowptr_[n] + n;
  rowptr_.swap(rp);
  colidx_.swap(ci);
  num_edges_ += n;
}

void LearningGraph::partition_rows(gai_host::Comm* comm, index_t nv_global) {
  partition_rows(comm->world(), comm->rank(), nv_global);
  comm_ = comm;
}

void LearningGraph::partition_rows(int world, int rank, index_t nv_global) {
  const gai_host::OwnerRange own = gai_host::owner_range(nv_global, world, rank);
  if (own.last - own.first != num_vertices_) {
    std::cerr << "partition_rows: this rank owns " << own.last - own.first << " vertices but the graph holds " << num_vertices_ << " rows\n";
    std::exit(EXIT_FAILURE);
  }
  nv_global_ = nv_global; first_ = own.first;
  // Distinct remote neighbours in ascending global id, and each edge's rank in that list. A bitmap over the global id space (14 MB for
  // the papers100M shape) + per-word prefix counts gives both in O(nnz + N/64): the sort / unique / binary-search form of the same
  // result took 30 s per rank at that size (88 M remote column ids), all of it set-up wall clock.
  const size_t words = ((size_t)nv_global + 63) / 64;
  std::vector<uint64_t> bits(words, 0);
  for (index_t c : colidx_)
    if (c < own.first || c >= own.last) bits[c >> 6] |= 1ull << (c & 63);
  std::vector<index_t> before(words + 1, 0);   // halo vertices with an id below word w
  for (size_t w = 0; w < words; w++) before[w + 1] = before[w] + (index_t)__builtin_popcountll(bits[w]);
  halo_gids_.assign(before[words], 0);
  for (size_t w = 0, k = 0; w < words; w++)
    for (uint64_t b = bits[w]; b; b &= b - 1) halo_gids_[k++] = (index_t)(w * 64 + (size_t)__builtin_ctzll(b));
  for (index_t& c : colidx_) {
    if (c >= own.first && c < own.last) c -= own.first;
    else c = num_vertices_ + before[c >> 6] + (index_t)__builtin_popcountll(bits[c >> 6] & ((1ull << (c & 63)) - 1));
  }
}

void LearningGraph::register_gather_buffer(const float* buf) {
  if (comm_) comm_->register_buffer(buf);
}

const float* LearningGraph::halo_exchange(const float* buf, int F, size_t ld) {
  if (!comm_ || comm_->world() == 1 || halo_gids_.empty()) {
    if (comm_ && comm_->world() > 1) halo_exchange_into(buf, F, ld, nullptr);  // a rank without halo still takes part in the barriers
    return nullptr;
  }
  ensure_halo_scratch(halo_gids_.size() * ld);
  halo_exchange_into(buf, F, ld, halo_scratch_);
  return halo_scratch_;
}

void LearningGraph::ensure_halo_scratch(size_t need) {
  if (need <= halo_scratch_floats_) return;  // grown on the first epoch only (widest exchanged matrix)
  if (halo_scratch_) {
    die_on(gai_stream_sync(gai_host::stream()), "gai_stream_sync");
    if (pull_stream_) die_on(gai_stream_sync(pull_stream_), "gai_stream_sync");
    gai_free(halo_scratch_);
  }
  void* p = nullptr;
  die_on(gai_malloc(&p, sizeof(float) * need), "gai_malloc(halo scratch)");
  halo_scratch_ = reinterpret_cast<float*>(p);
  halo_scratch_floats_ = need;
}

int LearningGraph::halo_block_count(int F) {
  const char* e = std::getenv("GAI_HALO_BLOCKS");   // read per call: the tests switch it between runs of one process
  int n = e ? std::atoi(e) : 1;
  n = n < 1 ? 1 : (n > 8 ? 8 : n);
  while (n > 1 && (F + n - 1) / n < 64) n--;   // narrower blocks waste gather lanes in the aggregation
  return n;
}

LearningGraph::HaloBlocks LearningGraph::halo_exchange_begin(const float* buf, int F, size_t ld) {
  HaloBlocks hb;
  if (!comm_ || comm_->world() == 1) return hb;
  const int id = comm_->id_of(buf);
  if (id < 0) { std::cerr << "halo_exchange: the gathered matrix was never registered with the peer group\n"; std::exit(EXIT_FAILURE); }
  hb.n = halo_block_count(F);
  const int w = ((F + hb.n - 1) / hb.n + 31) / 32 * 32;   // block width: a multiple of 32 columns (whole sign-bit words, 128-byte segments)
  hb.n = (F + w - 1) / w;
  for (int k = 0; k < hb.n; k++) { hb.col0[k] = k * w; hb.ncol[k] = (k + 1) * w <= F ? w : F - k * w; }
  if (!pull_stream_) {
    die_on(gai_stream_create(&pull_stream_), "gai_stream_create");
    die_on(gai_event_create(&ev_fork_), "gai_event_create");
    die_on(gai_event_create(&ev_done_), "gai_event_create");
    for (auto& e : ev_block_) die_on(gai_event_create(&e), "gai_event_create");
  }
  const bool have = !halo_gids_.empty();
  if (have) ensure_halo_scratch(halo_gids_.size() * ld);
  gai_stream_t main = gai_host::stream();
  {
    gai_host::OpScope sc("HALO", "barrier + fork W=" + std::to_string(F), 0, 0);
    die_on(gai_peers_barrier_on(comm_->peers(), 0, main), "gai_peers_barrier_on");   // every owner's matrix is complete
    die_on(gai_event_record(ev_fork_, main), "gai_event_record");
  }
  die_on(gai_stream_wait_event(pull_stream_, ev_fork_), "gai_stream_wait_event");
  for (int k = 0; k < hb.n; k++) {
    die_on(gai_halo_pull_cols(comm_->peers(), plan_, id, hb.col0[k], hb.ncol[k], ld, have ? halo_scratch_ : nullptr, ld,
                              GAI_PULL_NO_BARRIER_BEFORE | GAI_PULL_NO_BARRIER_AFTER | GAI_PULL_SMALL_GRID, 1, pull_stream_), "gai_halo_pull_cols");
    die_on(gai_event_record(ev_block_[k], pull_stream_), "gai_event_record");
  }
  die_on(gai_peers_barrier_on(comm_->peers(), 1, pull_stream_), "gai_peers_barrier_on");   // every rank has read what it needs
  die_on(gai_event_record(ev_done_, pull_stream_), "gai_event_record");
  halo_exchanges++;
  halo_bytes += 4ull * (unsigned long long)F * halo_gids_.size();
  hb.halo = have ? halo_scratch_ : nullptr;
  return hb;
}

void LearningGraph::halo_wait_block(int k) {
  if (!pull_stream_) return;
  gai_host::OpScope sc("HALO", "wait block", 0, 0);
  die_on(gai_stream_wait_event(gai_host::stream(), ev_block_[k]), "gai_stream_wait_event");
}

void LearningGraph::halo_exchange_end() {
  if (!pull_stream_) return;
  gai_host::OpScope sc("HALO", "wait closing barrier", 0, 0);
  die_on(gai_stream_wait_event(gai_host::stream(), ev_done_), "gai_stream_wait_event");
}

void LearningGraph::halo_exchange_into(const float* buf, int F, size_t ld, float* dst) {
  if (!comm_ || comm_->world() == 1) return;
  const int id = comm_->id_of(buf);
  if (id < 0) { std::cerr << "halo_exchange: the gathered matrix was never registered with the peer group\n"; std::exit(EXIT_FAILURE); }
  gai_host::OpScope sc("HALO", "pull W=" + std::to_string(F), 4.0 * F * halo_gids_.size(), 0);
  die_on(gai_halo_pull(comm_->peers(), plan_, id, F, ld, dst, ld, 0, gai_host::stream()), "gai_halo_pull");
  halo_exchanges++;
  halo_bytes += 4ull * (unsigned long long)F * halo_gids_.size();
}

void LearningGraph::compute_vertex_data() {
  if (!dev_) copy_to_gpu();  // norms are produced together with the device CSR
}
void LearningGraph::compute_edge_data() { compute_vertex_data(); }

void LearningGraph::dealloc() {
  if (dev_) { gai_csr_destroy(dev_); dev_ = nullptr; }
  if (plan_) { gai_halo_plan_destroy(plan_); plan_ = nullptr; }
  if (halo_scratch_) { gai_free(halo_scratch_); halo_scratch_ = nullptr; halo_scratch_floats_ = 0; }
  if (pull_stream_) {
    gai_stream_sync(pull_stream_);
    gai_event_destroy(ev_fork_); gai_event_destroy(ev_done_);
    for (auto& e : ev_block_) { gai_event_destroy(e); e = nullptr; }
    gai_stream_destroy(pull_stream_);
    pull_stream_ = nullptr; ev_fork_ = ev_done_ = nullptr;
  }
  rowptr_.clear(); rowptr_.shrink_to_fit();
  colidx_.clear(); colidx_.shrink_to_fit();
}

// ---- Reader -------------------------------------------------------------------------------------------------------

static const char* kDatasets[] = {"cora", "citeseer", "ppi", "pubmed", "flickr", "yelp", "reddit", "amazon", "tester",
                                  "ogbn-arxiv", "ogbn-products", "ogbn-proteins", "ogbn-papers100M"};  // include/gnn/configs.h:8-11

template <typename T>
static void read_exact(const std::string& fname, T* dst, size_t count) {
  std::ifstream in(fname.c_str(), std::ios::binary);
  if (!in.good()) { std::cerr << "Failed to open file: " << fname << "\n"; std::exit(1); }
  in.read(reinterpret_cast<char*>(dst), sizeof(T) * count);
}

// ---- legacy .csgr / text layout (reader.cpp:16-246) ----------------------------------------------------------------------------

static std::string dataset_root() {
  const char* root = std::getenv("DATASET_PATH");
  if (!root) { std::cerr << "DATASET_PATH is not set\n"; std::exit(1); }
  return std::string(root);
}

size_t Reader::csgr_read_labels(std::vector<label_t>& labels, bool is_single_class) {  // reader.cpp:16-64
  const std::string filename = dataset_root() + dataset_str + "/" + dataset_str + "-labels.txt";
  std::ifstream in(filename.c_str(), std::ios::in);
  std::string line;
  size_t m = 0, num_classes = 0;
  in >> m >> num_classes >> std::ws;
  gai_host::out() << (is_single_class ? "Using single-class (one-hot) labels\n" : "Using multi-class (multi-hot) labels\n");
  labels.assign(is_single_class ? m : m * num_classes, 0);
  gai_host::out() << "Number of classes (unique label counts): " << num_classes << "\n";
  size_t v = 0;
  while (std::getline(in, line) && v < m) {
    std::istringstream label_stream(line);
    unsigned x = 0;
    for (size_t idx = 0; idx < num_classes; ++idx) {
      label_stream >> x;
      if (is_single_class) {
        if (x != 0) { labels[v] = (label_t)idx; break; }  // the first non-zero column is the class
      } else {
        labels[v * num_classes + idx] = (label_t)x;
      }
    }
    v++;
  }
  num_vertex_classes = (int)num_classes;
  return num_classes;
}

size_t Reader::csgr_read_features(std::vector<float>& feats, std::string filetype) {  // reader.cpp:68-121
  size_t m = 0, flen = 0;
  const std::string base = dataset_root() + dataset_str + "/" + dataset_str;
  if (filetype == "bin") {
    std::ifstream dims((base + "-dims.txt").c_str(), std::ios::in);
    dims >> m >> flen >> std::ws;
    gai_host::out() << "Reading features ... N x D: " << m << " x " << flen << "\n";
    feats.assign(m * flen, 0.f);
    std::ifstream in((base + "-feats.bin").c_str(), std::ios::binary | std::ios::in);
    if (!feats.empty()) in.read(reinterpret_cast<char*>(feats.data()), sizeof(float) * feats.size());
  } else {  // "<row> <col> <value>" triples
    std::ifstream in((base + ".ft").c_str(), std::ios::in);
    in >> m >> flen >> std::ws;
    gai_host::out() << "Reading features ... N x D: " << m << " x " << flen << "\n";
    feats.assign(m * flen, 0.f);
    std::string line;
    while (std::getline(in, line)) {
      std::istringstream s(line);
      size_t u = 0, v = 0;
      float w = 0.f;
      s >> u >> v >> w;
      if (u < m && v < flen) feats[u * flen + v] = w;
    }
  }
  feat_len = (index_t)flen;
  return flen;
}

size_t Reader::csgr_read_masks(std::string mask_type, size_t n, size_t& begin, size_t& end, mask_t* masks) {  // reader.cpp:125-171
  bool known = false;
  for (const char* d : kDatasets) known = known || dataset_str == d;
  if (!known) { std::cout << "Dataset currently not supported\n"; std::exit(1); }
  const std::string filename = dataset_root() + dataset_str + "/" + dataset_str + "-" + mask_type + "_mask.txt";
  std::ifstream in(filename.c_str(), std::ios::in);
  std::string line;
  in >> begin >> end >> std::ws;
  size_t i = 0, sample_count = 0;
  while (std::getline(in, line)) {
    std::istringstream mask_stream(line);
    if (i >= begin && i < end) {
      unsigned mask = 0;
      mask_stream >> mask;
      if (mask == 1) { masks[i] = 1; sample_count++; }
    }
    i++;
  }
  gai_host::out() << mask_type << "_mask range: [" << begin << ", " << end << ") Number of valid samples: " << sample_count << " ("
                  << (float)sample_count / (float)n * 100.0f << "%)\n";
  return sample_count;
}

void Reader::csgr_read_graph(LearningGraph* g) {  // reader.cpp:173-246: Galois .gr v1 — header {version, sizeof(edge data), nv, ne}, u64 row ENDS, u32 columns
  gai_host::out() << "Reading graph into CPU memory\n";
  const std::string filename = dataset_root() + dataset_str + "/" + dataset_str + ".csgr";
  std::ifstream in(filename.c_str(), std::ios::binary);
  if (!in.good()) { std::cout << "LearningGraph: unable to open " << filename << "\n"; std::exit(1); }
  uint64_t header[4] = {0, 0, 0, 0};
  in.read(reinterpret_cast<char*>(header), sizeof(header));
  assert(header[0] == 1);
  if (header[1] != 0) { std::cout << "LearningGraph: currently edge data not supported.\n"; std::exit(1); }
  const uint64_t nv = header[2], ne = header[3];
  std::vector<uint64_t> ends(nv);
  in.read(reinterpret_cast<char*>(ends.data()), sizeof(uint64_t) * nv);
  g->allocateFrom((index_t)nv, (index_t)ne);
  in.read(reinterpret_cast<char*>(g->edge_dst_host_ptr()), sizeof(uint32_t) * ne);
  for (uint64_t v = 0; v < nv; v++) g->fixEndEdge((index_t)v, (index_t)ends[v]);
  const index_t* ci = g->edge_dst_host_ptr();
  for (uint64_t e = 0; e < ne; e++)
    if (ci[e] >= nv) { printf("\tinvalid edge to %u at index %llu.\n", ci[e], (unsigned long long)e); std::exit(0); }
  num_vertices_ = (index_t)nv; num_edges_ = (index_t)ne;
  g->degree_counting();
  gai_host::out() << "|V| " << nv << " |E| " << ne << " max_deg " << g->get_max_degree() << "\n";
}

void Reader::bin_read_graph(LearningGraph* g) {
  const char* root = std::getenv("DATASET_PATH");
  if (!root) { std::cerr << "DATASET_PATH is not set\n"; std::exit(1); }
  inputfile_path = std::string(root) + dataset_str + "/";
  gai_host::out() << "input file path: " << inputfile_path << ", graph name: " << dataset_str << "\n";
  std::ifstream meta((inputfile_path + "graph.meta.txt").c_str());
  if (!meta.good()) { std::cerr << "Failed to open file: " << inputfile_path << "graph.meta.txt\n"; std::exit(1); }
  int vid_size = 0, eid_size = 0, vlabel_size = 0, elabel_size = 0, max_degree = 0;
  int64_t nv = 0, ne = 0;
  meta >> nv >> ne >> vid_size >> eid_size >> vlabel_size >> elabel_size >> max_degree >> feat_len >> num_vertex_classes >> num_edge_classes;
  meta >> train_begin >> train_end >> train_count >> val_begin >> val_end >> val_count >> test_begin >> test_end >> test_count;
  assert(vid_size == 4 && eid_size == 8 && vlabel_size == 1);
  assert(max_degree > 0 && max_degree < nv);
  num_vertices_ = (index_t)nv;
  num_edges_ = (index_t)ne;
  g->allocateFrom(num_vertices_, num_edges_);
  std::vector<int64_t> rows((size_t)nv + 1);
  read_exact<int64_t>(inputfile_path + "graph.vertex.bin", rows.data(), rows.size());
  read_exact<index_t>(inputfile_path + "graph.edge.bin", g->edge_dst_host_ptr(), num_edges_);
  index_t* rp = g->row_start_host_ptr();
  for (size_t i = 0; i <= (size_t)nv; i++) rp[i] = (index_t)rows[i];  // on-disk int64 -> u32 (reader.cpp:445-454)
  g->degree_counting();
  gai_host::out() << "|V| " << nv << " |E| " << ne << " max_deg " << g->get_max_degree() << "\n";
}

size_t Reader::bin_read_features(std::vector<float>& feats) {
  gai_host::out() << "Reading features ... N x D: " << num_vertices_ << " x " << feat_len << "\n";
  feats.resize((size_t)num_vertices_ * feat_len);
  // as the reference (reader.cpp:258-263): no open check here — a dataset without a feature file (feat_len 0) loads with no features
  std::ifstream in((inputfile_path + "graph.feats.bin").c_str(), std::ios::binary);
  if (!feats.empty()) in.read(reinterpret_cast<char*>(feats.data()), sizeof(float) * feats.size());
  return feat_len;
}

int Reader::bin_read_vlabels(std::vector<label_t>& labels, bool is_single_class) {
  assert(num_vertex_classes > 0 && num_vertex_classes < 255);
  std::vector<vlabel_t> vl(num_vertices_, 0);
  std::ifstream probe((inputfile_path + "graph.vlabel.bin").c_str());
  gai_host::out() << (is_single_class ? "Using single-class (one-hot) labels\n" : "Using multi-class (multi-hot) labels\n");
  labels.assign(is_single_class ? (size_t)num_vertices_ : (size_t)num_vertices_ * num_vertex_classes, 0);
  if (probe.good()) {
    read_exact<vlabel_t>(inputfile_path + "graph.vlabel.bin", vl.data(), vl.size());
    for (size_t v = 0; v < num_vertices_; v++) {
      if (is_single_class) labels[v] = vl[v];
      else if (vl[v] < num_vertex_classes) labels[v * num_vertex_classes + vl[v]] = 1;
    }
  } else {
    // reader.cpp:386-408, quirks included: in single-class mode only the scratch array is drawn (labels stay all zero, so every
    // generated label is a valid class id); in multi-hot mode the drawn class is 1..C, and a draw of C sets no column at all
    gai_host::out() << "WARNING: vertex label file not exist; generating random labels\n";
    for (size_t v = 0; v < num_vertices_; v++) {
      const int rand_class = rand() % num_vertex_classes + 1;
      if (is_single_class) vl[v] = (vlabel_t)rand_class;
      else if (rand_class < num_vertex_classes) labels[v * num_vertex_classes + rand_class] = 1;
    }
  }
  gai_host::out() << "maximum vertex label: " << unsigned(*std::max_element(vl.begin(), vl.end())) << "\n";
  return num_vertex_classes;
}

size_t Reader::bin_read_masks(std::string mask_type, size_t n, size_t& begin, size_t& end, mask_t*) {
  bool known = false;
  for (const char* d : kDatasets) known = known || dataset_str == d;
  if (!known) { gai_host::out() << "Dataset currently not supported\n"; std::exit(1); }
  size_t count;
  if (mask_type == "train") { begin = train_begin; end = train_end; count = train_count; }
  else if (mask_type == "val") { begin = val_begin; end = val_end; count = val_count; }
  else { begin = test_begin; end = test_end; count = test_count; }
  gai_host::out() << mask_type << "_mask range: [" << begin << ", " << end << ") Number of valid samples: " << count << " ("
            << (float)count / (float)n * 100.0f << "%)\n";
  return count;  // the .masks.bin files are not read: ranges come from the meta file (reader.cpp:272-316)
}
