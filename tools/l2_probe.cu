// Measures the L2 -> SM read bandwidth cap of this GPU (the LTS throughput cap the aggregation kernels run into): every SM streams a
// buffer that fits the L2 (default 48 MB) with 128-bit loads that bypass L1 (ld.global.cg), many passes; prints GB/s per pass count
// and the same for a buffer far larger than L2 (DRAM-bound reference point). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// tools/l2_probe.cu -o tools/_bin/l2_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) read_kernel(const float4* __restrict__ p, size_t n4, int passes, float* sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < passes; it++) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) v[u] = __ldcg(p + i + u * stride);
#pragma unroll
      for (int u = 0; u < 4; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { const float4 v = __ldcg(p + i); acc += v.x + v.y + v.z + v.w; }
  }
  if (acc == 123456.789f) *sink = acc;
}

int main(int argc, char** argv) {
  int sms = 0, mhz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  float* sink; cudaMalloc(&sink, 4);
  const size_t sizes_mb[] = {16, 32, 48, 64, 96, 2048};
  for (size_t mb : sizes_mb) {
    const size_t bytes = mb << 20, n4 = bytes / 16;
    float4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    const int passes = mb >= 1024 ? 4 : 64;
    for (int ctas_per_sm = 2; ctas_per_sm <= 4; ctas_per_sm += 2) {
      read_kernel<<<sms * ctas_per_sm, 512>>>(buf, n4, 2, sink);  // warm the L2
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      read_kernel<<<sms * ctas_per_sm, 512>>>(buf, n4, passes, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      const double gbs = (double)bytes * passes / (ms * 1e-3) / 1e9;
      printf("{\"buffer_MB\": %zu, \"ctas_per_sm\": %d, \"passes\": %d, \"ms\": %.4f, \"read_GBps\": %.1f, \"bytes_per_clk_at_max_clock\": %.0f, \"sms\": %d, \"max_clock_khz\": %d}\n",
             mb, ctas_per_sm, passes, ms, gbs, gbs * 1e9 / (mhz * 1e3), sms, mhz);
    }
    cudaFree(buf);
  }
  return 0;
}
