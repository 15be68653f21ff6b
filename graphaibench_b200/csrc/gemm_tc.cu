// tcgen05 / TMEM dense transform (3xTF32). Placeholder translation unit: the kernel lands in a later commit;
// until then every shape is declined and gai_matmul uses the fp32 SIMT path.
#include "gai_internal.cuh"
namespace gai {
int gemm_tc(size_t, size_t, size_t, const float*, size_t, const float*, size_t, float*, size_t, int, int, int, int, int, cudaStream_t) {
  return GAI_ERR_UNSUPPORTED;
}
}  // namespace gai
