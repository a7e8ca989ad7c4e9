// C-ABI glue: version, thread-local error text, GEMM dispatch between the fp32 CUDA-core kernel
// (gemm_simt.cu) and the tcgen05/TMA kernel (gemm_tcgen05.cu).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

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

// ---- host side of a batch (replaces the per-batch numpy copies of GraphTokenBlockDataset.__getitem__ + collater)
// dst <- src (bytes), split over n_threads host threads; check_i64: the range holds int64 ids, every one of which must lie in
// [lo, hi) -- *bad is set to 1 otherwise (the copy still completes).  dst == NULL: check only (page-locked sources are not copied).
// The per-batch host work of GraphTokenBlockDataset.__getitem__ (neighbor_offsets[offsets], precompute_feats[offsets],
// token_block_dataset.py:309,327-329) as ONE blocking native call without the interpreter lock.
extern "C" int32_t gnnlm_host_copy(void* dst, const void* src, int64_t bytes, int32_t check_i64, int64_t lo, int64_t hi,
                                   int32_t n_threads, int32_t* bad) {
  GNNLM_CHECK_ARG(src && bytes >= 0 && (dst || check_i64), GNNLM_E_ARG, "gnnlm_host_copy: null pointer / bad size");
  GNNLM_CHECK_ARG(!check_i64 || (bytes % 8 == 0 && (uintptr_t)src % 8 == 0 && bad), GNNLM_E_ARG,
                  "gnnlm_host_copy: an id check needs 8 B aligned int64 data and a flag");
  if (bytes == 0) return 0;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 32) n_threads = 32;
  const int64_t min_chunk = 1 << 20;                              // below ~1 MB per thread a spawn costs more than it saves
  int64_t nt = bytes / min_chunk;
  if (nt < 1) nt = 1;
  if (nt > n_threads) nt = n_threads;
  std::atomic<int> any_bad(0);
  auto work = [&](int64_t b0, int64_t b1) {
    if (dst) memcpy(reinterpret_cast<char*>(dst) + b0, reinterpret_cast<const char*>(src) + b0, (size_t)(b1 - b0));
    if (check_i64) {
      const int64_t* p = reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(src) + b0);
      const int64_t n = (b1 - b0) / 8;
      int64_t mn = lo, mx = lo;
      for (int64_t i = 0; i < n; ++i) {
        const int64_t v = p[i];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
      }
      if (mn < lo || mx >= hi) any_bad.store(1);
    }
  };
  const int64_t per = ((bytes / nt) + 63) / 64 * 64;              // 64 B aligned pieces (multiples of 8 B)
  std::vector<std::thread> pool;
  for (int64_t t = 1; t < nt; ++t) {
    const int64_t b0 = t * per, b1 = (t + 1 == nt) ? bytes : (t + 1) * per;
    if (b0 < bytes) pool.emplace_back(work, b0, b1 < bytes ? b1 : bytes);
  }
  work(0, per < bytes ? per : bytes);
  for (auto& th : pool) th.join();
  if (check_i64 && any_bad.load()) *bad = 1;
  return 0;
}
