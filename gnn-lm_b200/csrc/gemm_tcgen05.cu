// placeholder until the tcgen05 kernel lands (next commit): reports "unsupported" loudly.
#include "common.cuh"
namespace gnnlm {
int32_t gemm_tc_supported() { return 0; }
int64_t gemm_tc_lse_tile_n() { return 128; }
int32_t gemm_tc_store(const void*, int32_t, int64_t, const void*, const void*, int64_t, const float*, const float*, int64_t,
                      void*, int32_t, int64_t, int64_t, const int32_t*, int64_t, int64_t, int32_t, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return GNNLM_E_UNSUPPORTED;
}
int32_t gemm_tc_lse(const void*, int32_t, int64_t, const void*, const void*, int64_t, const int32_t*, float*, float*, float*,
                    int64_t, const int32_t*, int64_t, int64_t, int32_t, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return GNNLM_E_UNSUPPORTED;
}
}  // namespace gnnlm
