// C-ABI glue: version, thread-local error text, GEMM dispatch between the fp32 CUDA-core kernel
// (gemm_simt.cu) and the tcgen05/TMA kernel (gemm_tcgen05.cu).
#include <stdarg.h>

#include "common.cuh"

namespace gnnlm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int32_t gemm_simt_store(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                        const float* residual, int64_t ldr, void* C, int32_t c_dtype, int64_t ldc, int64_t M,
                        const int32_t* m_dev, int64_t N, int64_t K, cudaStream_t st);
int32_t gemm_simt_lse(const float* A, int64_t lda, const float* W, int64_t ldw, const int32_t* pick, float* part_max,
                      float* part_sum, float* picked, int64_t M, const int32_t* m_dev, int64_t N, int64_t K,
                      cudaStream_t st);
// gemm_tcgen05.cu
int32_t gemm_tc_supported();
int64_t gemm_tc_lse_tile_n();
int32_t gemm_tc_store(const void* A, int32_t a_dtype, int64_t lda, const void* W, const void* W_lo, float w_scale, int64_t ldw,
                      const float* bias, const void* residual, int32_t r_dtype, int64_t ldr, void* C, int32_t c_dtype,
                      int64_t ldc, int64_t M, const int32_t* m_dev, int64_t N, int64_t K, int32_t math, cudaStream_t st);
int32_t gemm_tc_lse(const void* A, int32_t a_dtype, int64_t lda, const void* W, const void* W_lo, float w_scale, int64_t ldw,
                    const int32_t* pick, float* part_max, float* part_sum, float* picked, int64_t M,
                    const int32_t* m_dev, int64_t N, int64_t K, int32_t math, cudaStream_t st);

int32_t gemm_tc_batched_f16x3(const void* A, int64_t lda, int64_t a_bs, const void* W_hi, const void* W_lo, int64_t ldw,
                              int64_t w_bs, float w_scale, const float* residual, int64_t ldr, int64_t r_bs, float* C,
                              int64_t ldc, int64_t c_bs, int64_t nb, int64_t M, int64_t N, int64_t K, int32_t causal,
                              cudaStream_t st);

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_version(void) { return 100; }
extern "C" const char* gnnlm_last_error(void) { return g_err; }
extern "C" int32_t gnnlm_has_tcgen05(void) { return gemm_tc_supported(); }

extern "C" int64_t gnnlm_lse_num_tiles(int64_t N, int32_t math) {
  const int64_t tn = math == GNNLM_MATH_FP32_SIMT ? 128 : gemm_tc_lse_tile_n();
  return (N + tn - 1) / tn;
}

static int32_t check_linear(const char* who, const void* A, int32_t a_dtype, const void* W, int64_t lda, int64_t ldw,
                            int64_t M, int64_t N, int64_t K, int32_t math) {
  GNNLM_CHECK_ARG(A && W, GNNLM_E_ARG, "%s: null operand", who);
  GNNLM_CHECK_ARG(a_dtype != GNNLM_F16X2 || math == GNNLM_MATH_F16X3, GNNLM_E_UNSUPPORTED,
                  "%s: split-fp16 operands are the MATH_F16X3 activation format", who);
  GNNLM_CHECK_ARG(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K, GNNLM_E_SHAPE, "%s: bad shape M=%lld N=%lld K=%lld lda=%lld ldw=%lld",
                  who, (long long)M, (long long)N, (long long)K, (long long)lda, (long long)ldw);
  GNNLM_CHECK_ARG(math >= GNNLM_MATH_FP32_SIMT && math <= GNNLM_MATH_F16X3, GNNLM_E_ARG, "%s: unknown math mode %d", who, math);
  if (math == GNNLM_MATH_BF16)
    GNNLM_CHECK_ARG(a_dtype == GNNLM_BF16, GNNLM_E_UNSUPPORTED, "%s: MATH_BF16 needs bf16 operands", who);
  else if (math == GNNLM_MATH_F16X3)
    GNNLM_CHECK_ARG(a_dtype == GNNLM_F32 || a_dtype == GNNLM_F16X2, GNNLM_E_UNSUPPORTED,
                    "%s: MATH_F16X3 takes fp32 or split-fp16 A", who);
  else
    GNNLM_CHECK_ARG(a_dtype == GNNLM_F32, GNNLM_E_UNSUPPORTED, "%s: fp32/tf32 math needs fp32 operands", who);
  return 0;
}

extern "C" int32_t gnnlm_linear(const void* A, int32_t a_dtype, int64_t lda, const void* W, const void* W_lo, float w_scale,
                                int64_t ldw,
                                const float* bias, const void* residual, int32_t r_dtype, int64_t ldr, void* C,
                                int32_t c_dtype, int64_t ldc, int64_t M, const int32_t* m_dev, int64_t N, int64_t K,
                                int32_t math, gnnlm_stream_t stream) {
  int32_t rc = check_linear("gnnlm_linear", A, a_dtype, W, lda, ldw, M, N, K, math);
  if (rc) return rc;
  GNNLM_CHECK_ARG(C && ldc >= N * (c_dtype == GNNLM_F16X2 ? 2 : 1), GNNLM_E_ARG, "gnnlm_linear: bad C/ldc");
  GNNLM_CHECK_ARG(c_dtype == GNNLM_F32 || c_dtype == GNNLM_BF16 || (c_dtype == GNNLM_F16X2 && math != GNNLM_MATH_FP32_SIMT),
                  GNNLM_E_UNSUPPORTED, "gnnlm_linear: C dtype");
  GNNLM_CHECK_ARG(!residual || ldr >= N * (r_dtype == GNNLM_F16X2 ? 2 : 1), GNNLM_E_SHAPE, "gnnlm_linear: ldr too small");
  GNNLM_CHECK_ARG(!residual || r_dtype == GNNLM_F32 ||
                      ((r_dtype == GNNLM_BF16 || r_dtype == GNNLM_F16X2) && math != GNNLM_MATH_FP32_SIMT),
                  GNNLM_E_UNSUPPORTED, "gnnlm_linear: residual must be F32 (or BF16 / F16X2 on the tensor-core path)");
  if (math == GNNLM_MATH_FP32_SIMT)
    return gemm_simt_store((const float*)A, lda, (const float*)W, ldw, bias, (const float*)residual, ldr, C, c_dtype, ldc, M,
                           m_dev, N, K, (cudaStream_t)stream);
  return gemm_tc_store(A, a_dtype, lda, W, W_lo, w_scale, ldw, bias, residual, r_dtype, ldr, C, c_dtype, ldc, M, m_dev, N, K,
                       math, (cudaStream_t)stream);
}

extern "C" int32_t gnnlm_linear_lse(const void* A, int32_t a_dtype, int64_t lda, const void* W, const void* W_lo,
                                    float w_scale, int64_t ldw, const int32_t* pick, float* part_max, float* part_sum, float* picked,
                                    int64_t M, const int32_t* m_dev, int64_t N, int64_t K, int32_t math,
                                    gnnlm_stream_t stream) {
  int32_t rc = check_linear("gnnlm_linear_lse", A, a_dtype, W, lda, ldw, M, N, K, math);
  if (rc) return rc;
  GNNLM_CHECK_ARG(part_max && part_sum && picked, GNNLM_E_ARG, "gnnlm_linear_lse: null output");
  if (math == GNNLM_MATH_FP32_SIMT)
    return gemm_simt_lse((const float*)A, lda, (const float*)W, ldw, pick, part_max, part_sum, picked, M, m_dev, N, K,
                         (cudaStream_t)stream);
  return gemm_tc_lse(A, a_dtype, lda, W, W_lo, w_scale, ldw, pick, part_max, part_sum, picked, M, m_dev, N, K, math,
                     (cudaStream_t)stream);
}

extern "C" int32_t gnnlm_linear_batched_f16x3(const void* A, int64_t lda, int64_t a_bs, const void* W_hi, const void* W_lo,
                                              int64_t ldw, int64_t w_bs, float w_scale, const float* residual, int64_t ldr,
                                              int64_t r_bs, float* C, int64_t ldc, int64_t c_bs, int64_t nb, int64_t M,
                                              int64_t N, int64_t K, int32_t causal, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(A && W_hi && W_lo && C, GNNLM_E_ARG, "gnnlm_linear_batched_f16x3: null pointer");
  GNNLM_CHECK_ARG(causal >= 0 && causal <= 2, GNNLM_E_ARG, "gnnlm_linear_batched_f16x3: causal must be 0, 1 or 2");
  GNNLM_CHECK_ARG(nb >= 1 && M >= 0 && N > 0 && K > 0 && ldc >= N && (!residual || ldr >= N) && w_scale > 0.f, GNNLM_E_SHAPE,
                  "gnnlm_linear_batched_f16x3: bad shape");
  return gemm_tc_batched_f16x3(A, lda, a_bs, W_hi, W_lo, ldw, w_bs, w_scale, residual, ldr, r_bs, C, ldc, c_bs, nb, M, N, K,
                               causal, (cudaStream_t)stream);
}
