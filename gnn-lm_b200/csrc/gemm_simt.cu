// fp32 CUDA-core GEMM: the strict-parity / validation arithmetic mode (GNNLM_MATH_FP32_SIMT).
//
// C[m,n] = sum_k A[m,k] * W[n,k] (+ bias[n]) (+ residual[m,n])   -- nn.Linear layout, both operands
// K-contiguous.  This is NOT the performance path (gemm_tcgen05.cu is); it exists so that (a) every
// other kernel can be validated on the GPU independently of the tensor-core kernel, and (b) users can
// ask for bit-for-bit fp32 FMA accumulation like the reference's fp32 cuBLAS path
// (reference call sites: fairseq/models/hgt.py:320-322,347-348,401; knn/pq_wrapper.py:202;
// fairseq/modules/adaptive_softmax.py:184,197,202).
//
// Tiling: 128x128x16 per CTA, 256 threads, 8x8 register micro-tile, smem double buffering.
// Two epilogues: store (bias/residual fused) and row log-sum-exp partials + column pick
// (the [rows, vocab] tensor is never written).
#include "common.cuh"

namespace gnnlm {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;
constexpr int PAD = 4;

struct SimtEpiStore {
  const float* bias;
  const float* residual;
  int64_t ldr;
  void* C;
  int64_t ldc;
  int c_dtype;
};
struct SimtEpiLse {
  const int32_t* pick;
  float* part_max;
  float* part_sum;
  float* picked;
  int64_t n_tiles;
};

template <bool LSE, bool VEC>
__global__ void __launch_bounds__(GT) gemm_simt_kernel(const float* __restrict__ A, int64_t lda,
                                                       const float* __restrict__ W, int64_t ldw, int64_t M_cap,
                                                       const int32_t* __restrict__ m_dev, int64_t N, int64_t K,
                                                       SimtEpiStore es, SimtEpiLse el) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[2][BK][BN + PAD];
  const int64_t M = live_rows(M_cap, m_dev);
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  if (m0 >= M) return;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;        // 16 x 16 threads; thread owns rows ty*8.., cols tx*8..
  // global->smem: each thread loads 2 float4 of A and 2 of W per k-tile
  const int lrow = tid >> 2;                     // 0..63 (+64 for the second load)
  const int lk = (tid & 3) * 4;                  // 0,4,8,12
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rw[2];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int64_t m = m0 + lrow + h * 64, n = n0 + lrow + h * 64;
      if constexpr (VEC) {
        ra[h] = (m < M) ? __ldg(reinterpret_cast<const float4*>(A + m * lda + k0 + lk)) : make_float4(0, 0, 0, 0);
        rw[h] = (n < N) ? __ldg(reinterpret_cast<const float4*>(W + n * ldw + k0 + lk)) : make_float4(0, 0, 0, 0);
      } else {
        float t[4], u[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int64_t kk = k0 + lk + e;
          t[e] = (m < M && kk < K) ? __ldg(A + m * lda + kk) : 0.f;
          u[e] = (n < N && kk < K) ? __ldg(W + n * ldw + kk) : 0.f;
        }
        ra[h] = make_float4(t[0], t[1], t[2], t[3]);
        rw[h] = make_float4(u[0], u[1], u[2], u[3]);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y; Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  const int64_t nk = (K + BK - 1) / BK;
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 8]);
      float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  if constexpr (!LSE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + ty * 8 + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t n = n0 + tx * 8 + j;
        if (n >= N) continue;
        float v = acc[i][j];
        if (es.bias) v += __ldg(es.bias + n);
        if (es.residual) v += __ldg(es.residual + m * es.ldr + n);
        if (es.c_dtype == GNNLM_F32) reinterpret_cast<float*>(es.C)[m * es.ldc + n] = v;
        else reinterpret_cast<__nv_bfloat16*>(es.C)[m * es.ldc + n] = __float2bfloat16(v);
      }
    }
  } else {
    // per row: max / sum-exp over this tile's valid columns, reduced over the 16 threads (tx) that
    // share the row -- they are 16 consecutive lanes, so xor-shuffles 8,4,2,1 stay inside the group.
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + ty * 8 + i;
      const int32_t want = (m < M && el.pick) ? __ldg(el.pick + m) : -1;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t n = n0 + tx * 8 + j;
        if (n < N) {
          mx = fmaxf(mx, acc[i][j]);
          if (n == want) el.picked[m] = acc[i][j];
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t n = n0 + tx * 8 + j;
        if (n < N) s += expf(acc[i][j] - mx);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tx == 0 && m < M) {
        el.part_max[m * el.n_tiles + blockIdx.y] = mx;
        el.part_sum[m * el.n_tiles + blockIdx.y] = s;
      }
    }
  }
}

int32_t gemm_simt_store(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                        const float* residual, int64_t ldr, void* C, int32_t c_dtype, int64_t ldc, int64_t M,
                        const int32_t* m_dev, int64_t N, int64_t K, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  const bool vec = K % BK == 0 && lda % 4 == 0 && ldw % 4 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)W % 16 == 0;
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN));
  SimtEpiStore es{bias, residual, ldr, C, ldc, c_dtype};
  SimtEpiLse el{};
  if (vec) gemm_simt_kernel<false, true><<<grid, GT, 0, st>>>(A, lda, W, ldw, M, m_dev, N, K, es, el);
  else gemm_simt_kernel<false, false><<<grid, GT, 0, st>>>(A, lda, W, ldw, M, m_dev, N, K, es, el);
  GNNLM_LAUNCH_CHECK("gemm_simt_store");
  return 0;
}

int32_t gemm_simt_lse(const float* A, int64_t lda, const float* W, int64_t ldw, const int32_t* pick, float* part_max,
                      float* part_sum, float* picked, int64_t M, const int32_t* m_dev, int64_t N, int64_t K,
                      cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  const bool vec = K % BK == 0 && lda % 4 == 0 && ldw % 4 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)W % 16 == 0;
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN));
  SimtEpiStore es{};
  SimtEpiLse el{pick, part_max, part_sum, picked, ceil_div(N, BN)};
  if (vec) gemm_simt_kernel<true, true><<<grid, GT, 0, st>>>(A, lda, W, ldw, M, m_dev, N, K, es, el);
  else gemm_simt_kernel<true, false><<<grid, GT, 0, st>>>(A, lda, W, ldw, M, m_dev, N, K, es, el);
  GNNLM_LAUNCH_CHECK("gemm_simt_lse");
  return 0;
}

}  // namespace gnnlm
