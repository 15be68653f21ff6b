// Runtime glue behind include/gai_b200.h: status/error text, device memory, transfers, workspace.
// Replaces the reference's float_malloc_device / copy_*_device helpers (src/utilities/math_functions.cu:56-111)
// and its exit-on-error macros (include/utils/cutils.h:133-174) with int status codes.
#include <mutex>
#include <map>
#include <tuple>
#include "gai_internal.cuh"

namespace gai {

int sm_count();
static thread_local std::string g_err;
unsigned long long g_launches = 0;

int set_error(int code, const char* what, const char* detail) {
  g_err = std::string(what ? what : "") + ": " + (detail ? detail : "");
  return code;
}

uint32_t hub_degree_for(uint64_t nnz) {
  const uint64_t slots = (uint64_t)sm_count() * 32;
  uint64_t t = nnz / (slots ? slots : 1);
  if (t < 1024) t = 1024;
  if (t > 8192) t = 8192;
  return (uint32_t)t;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

struct Ws { void* p = nullptr; size_t bytes = 0; };
static std::mutex g_ws_mu;
static std::map<std::tuple<int, cudaStream_t, int>, Ws> g_ws;

int workspace(size_t bytes, void** out, cudaStream_t st) { return workspace_slot(0, bytes, out, st); }

int workspace_slot(int slot, size_t bytes, void** out, cudaStream_t st) {
  int dev = 0;
  GAI_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_ws_mu);
  Ws& w = g_ws[std::make_tuple(dev, st, slot)];
  if (w.bytes < bytes) {
    if (w.p) {
      GAI_CUDA(cudaStreamSynchronize(st));  // only work on this stream can still be reading the old buffer
      GAI_CUDA(cudaFree(w.p));
      w.p = nullptr; w.bytes = 0;
    }
    size_t want = bytes < (size_t(64) << 20) ? (size_t(64) << 20) : bytes;
    GAI_CUDA(cudaMalloc(&w.p, want));
    w.bytes = want;
  }
  *out = w.p;
  return GAI_OK;
}

}  // namespace gai

extern "C" {

const char* gai_last_error(void) { return gai::g_err.c_str(); }
int gai_version(void) { return 100; }
uint64_t gai_launch_count(void) { return __atomic_load_n(&gai::g_launches, __ATOMIC_RELAXED); }
int gai_event_create(void** ev) {
  GAI_CHECK_ARG(ev != nullptr);
  cudaEvent_t e;
  GAI_CUDA(cudaEventCreate(&e));
  *ev = e;
  return GAI_OK;
}
int gai_event_record(void* ev, gai_stream_t s) { GAI_CUDA(cudaEventRecord((cudaEvent_t)ev, gai::S(s))); return GAI_OK; }
int gai_event_elapsed_ms(void* a, void* b, float* ms) {
  GAI_CHECK_ARG(ms != nullptr);
  GAI_CUDA(cudaEventSynchronize((cudaEvent_t)b));
  GAI_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return GAI_OK;
}
int gai_event_destroy(void* ev) { if (ev) GAI_CUDA(cudaEventDestroy((cudaEvent_t)ev)); return GAI_OK; }

int gai_device_count(int* n) {
  GAI_CHECK_ARG(n != nullptr);
  GAI_CUDA(cudaGetDeviceCount(n));
  return GAI_OK;
}
int gai_set_device(int dev) { GAI_CUDA(cudaSetDevice(dev)); return GAI_OK; }

int gai_malloc(void** p, size_t bytes) {
  GAI_CHECK_ARG(p != nullptr);
  *p = nullptr;
  if (bytes == 0) return GAI_OK;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return gai::set_error(GAI_ERR_NOMEM, "cudaMalloc", cudaGetErrorString(e)); }
  GAI_CUDA(e);
  return GAI_OK;
}
int gai_free(void* p) { if (p) GAI_CUDA(cudaFree(p)); return GAI_OK; }
int gai_memset(void* p, int value, size_t bytes, gai_stream_t s) { if (bytes) GAI_CUDA(cudaMemsetAsync(p, value, bytes, gai::S(s))); return GAI_OK; }
int gai_memcpy_h2d(void* dst, const void* src, size_t bytes, gai_stream_t s) { if (bytes) GAI_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, gai::S(s))); return GAI_OK; }
int gai_memcpy_d2h(void* dst, const void* src, size_t bytes, gai_stream_t s) { if (bytes) GAI_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, gai::S(s))); return GAI_OK; }
int gai_memcpy2d(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes, size_t width_bytes, size_t rows, gai_stream_t s) {
  if (width_bytes && rows) GAI_CUDA(cudaMemcpy2DAsync(dst, dst_pitch_bytes, src, src_pitch_bytes, width_bytes, rows, cudaMemcpyDefault, gai::S(s)));
  return GAI_OK;
}
int gai_memcpy_d2d(void* dst, const void* src, size_t bytes, gai_stream_t s) { if (bytes) GAI_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, gai::S(s))); return GAI_OK; }
int gai_stream_sync(gai_stream_t s) { GAI_CUDA(cudaStreamSynchronize(gai::S(s))); return GAI_OK; }
int gai_stream_create(gai_stream_t* s) {
  GAI_CHECK_ARG(s != nullptr);
  cudaStream_t st = nullptr;
  GAI_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = reinterpret_cast<gai_stream_t>(st);
  return GAI_OK;
}
int gai_stream_destroy(gai_stream_t s) { if (s) GAI_CUDA(cudaStreamDestroy(gai::S(s))); return GAI_OK; }
int gai_stream_wait_event(gai_stream_t s, void* ev) {
  GAI_CHECK_ARG(ev != nullptr);
  GAI_CUDA(cudaStreamWaitEvent(gai::S(s), reinterpret_cast<cudaEvent_t>(ev), 0));
  return GAI_OK;
}
int gai_host_alloc_pinned(void** p, size_t bytes) { GAI_CHECK_ARG(p != nullptr); GAI_CUDA(cudaMallocHost(p, bytes ? bytes : 1)); return GAI_OK; }
int gai_host_free_pinned(void* p) { if (p) GAI_CUDA(cudaFreeHost(p)); return GAI_OK; }

}  // extern "C"
