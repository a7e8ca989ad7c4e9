// kNN-LM similarity recompute -- metric_type `l2` / `ip` of KNNModel.get_knn_prob (reference:
// knn/knn_model.py:159-177): sims[t, j] = -||q_t - key[ids[t,j]]||^2 or <q_t, key[ids[t,j]]> (keys L2-normalised first
// for a cosine index, :171-172; queries normalised for it at :181-184).  find_knn.py:65-66 keeps only the neighbour
// ids, so this is what turns the precomputed ids back into similarities without faiss.
//
//   knn_sims_keys_kernel  keys are the datastore's own fp16 / fp32 rows ([N_d, d] in HBM).  One CTA per token, the query
//                         in shared memory, one warp per neighbour with KS_ROWS rows in flight; HBM-bound row gathers:
//                         T * k_nn * d * sizeof(key) bytes.
//   knn_sims_pq_kernel    keys are only available PQ-compressed (codes [N_d, M] uint8, the 13 GB wiki103 table that is
//                         already resident for the graph): asymmetric distance computation.  With x^ = (y - b) A the
//                         decoded key (knn/pq_wrapper.py:169-203; y = concatenated centroids),
//                             <q, x^>       = sum_m <q'_m, cen[m, c_m]> - <q', b>,                     q' = q A^T
//                             ||q - x^||^2  = ||q||^2 - ||q'||^2 + sum_m ||q'_m + b_m - cen[m, c_m]||^2   (A A^T = I)
//                         so each token builds an [M, 256] table in shared memory once and every neighbour costs M
//                         code bytes + M table lookups instead of d key elements.
#include "common.cuh"

namespace gnnlm {

constexpr int KS_THREADS = 512;
constexpr int KS_ROWS = 4;          // neighbour rows in flight per warp

template <typename KT>
__device__ __forceinline__ void load8(const KT* p, float (&r)[8]) {
  if constexpr (sizeof(KT) == 2) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      r[2 * i] = f.x; r[2 * i + 1] = f.y;
    }
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  }
}

// metric: 0 = l2 (-sum (q-k)^2), 1 = ip.  normalise: bit 0 = keys (cosine `ip`), bit 1 = queries (cosine index).
template <typename KT>
__global__ void __launch_bounds__(KS_THREADS) knn_sims_keys_kernel(const float* __restrict__ q, int64_t ldq,
                                                                   const KT* __restrict__ keys, int64_t n_datastore, int d,
                                                                   const int64_t* __restrict__ ids, int k_nn, int metric,
                                                                   int normalise, float* __restrict__ sims) {
  extern __shared__ float sq[];                                  // d floats + 16 partials
  __shared__ float s_inv;
  const int64_t t = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = KS_THREADS / 32;
  float part = 0.f;
  for (int i = threadIdx.x; i < d; i += KS_THREADS) {
    const float x = __ldg(q + t * ldq + i);
    sq[i] = x;
    part = fmaf(x, x, part);
  }
  part = warp_sum(part);
  if (lane == 0) sq[d + warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < n_warps; ++w) s += sq[d + w];
    s_inv = (normalise & 2) ? 1.f / sqrtf(s) : 1.f;
  }
  __syncthreads();
  const float q_inv = s_inv;
  for (int j0 = warp * KS_ROWS; j0 < k_nn; j0 += n_warps * KS_ROWS) {
    const KT* row[KS_ROWS];
#pragma unroll
    for (int r = 0; r < KS_ROWS; ++r) {
      int64_t id = j0 + r < k_nn ? __ldg(ids + t * k_nn + j0 + r) : 0;
      if (id < 0) id += n_datastore;                            // numpy wrap of -1 (knn_model.py:163); masked by the caller
      if (id < 0 || id >= n_datastore) id = 0;
      row[r] = keys + id * (int64_t)d;
    }
    float acc[KS_ROWS], nrm[KS_ROWS];
#pragma unroll
    for (int r = 0; r < KS_ROWS; ++r) acc[r] = nrm[r] = 0.f;
    for (int c = lane * 8; c < d; c += 256) {
      float kk[KS_ROWS][8];
#pragma unroll
      for (int r = 0; r < KS_ROWS; ++r) load8<KT>(row[r] + c, kk[r]);
#pragma unroll
      for (int r = 0; r < KS_ROWS; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float qq = sq[c + i] * q_inv, x = kk[r][i];
          if (metric == 0) {
            const float df = qq - x;
            acc[r] = fmaf(df, df, acc[r]);
          } else {
            acc[r] = fmaf(qq, x, acc[r]);
            nrm[r] = fmaf(x, x, nrm[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < KS_ROWS; ++r) {
      const float a = warp_sum(acc[r]);
      const float n2 = (normalise & 1) ? warp_sum(nrm[r]) : 1.f;
      if (lane == 0 && j0 + r < k_nn)
        sims[t * k_nn + j0 + r] = metric == 0 ? -a : ((normalise & 1) ? a / sqrtf(n2) : a);
    }
  }
}

// qr = q A^T (the caller's GEMM) [T, M*dsub]; table[m][c] per token in shared memory.
// normalise (cosine index, knn_model.py:171-172,181-184): bit 1 = score with q / ||q|| (the rotation is linear, so qr is scaled
// by the same factor); bit 0 (ip only) = divide by ||x^|| = sqrt(sum_m key_norm2[m][c_m]), key_norm2[m][c] = ||cen[m,c] - b_m||^2
// (A A^T = I).
__global__ void __launch_bounds__(1024) knn_sims_pq_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ qr,
                                                           int64_t ldqr, int d_q, const uint8_t* __restrict__ codes,
                                                           int64_t n_datastore, int M, int dsub,
                                                           const float* __restrict__ cen, const float* __restrict__ bias,
                                                           const int64_t* __restrict__ ids, int k_nn, int metric,
                                                           const float* __restrict__ key_norm2, int normalise,
                                                           float* __restrict__ sims) {
  extern __shared__ float smem[];
  float* table = smem;                                           // [M, 256]
  float* sqr = smem + (size_t)M * 256;                           // [M*dsub] rotated query (+ b for l2)
  __shared__ float s_red[32];
  __shared__ float s_const, s_qn2;
  const int64_t t = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int dr = M * dsub;
  float part = 0.f;
  if (metric == 0 || (normalise & 2)) {                          // ||q||^2
    for (int i = threadIdx.x; i < d_q; i += blockDim.x) {
      const float x = __ldg(q + t * ldq + i);
      part = fmaf(x, x, part);
    }
    part = warp_sum(part);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < n_warps; ++w) s += s_red[w];
      s_qn2 = s;
    }
    __syncthreads();
  }
  const float q_inv = (normalise & 2) ? rsqrtf(s_qn2) : 1.f;
  // constant term: ip: -<q', b>;  l2: ||q||^2 - ||q'||^2
  part = 0.f;
  for (int i = threadIdx.x; i < dr; i += blockDim.x) {
    const float x = __ldg(qr + t * ldqr + i) * q_inv, b = bias ? __ldg(bias + i) : 0.f;
    sqr[i] = metric == 0 ? x + b : x;
    part += metric == 0 ? -x * x : -x * b;
  }
  part = warp_sum(part);
  __syncthreads();                                               // s_red is reused
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = metric == 0 ? s_qn2 * q_inv * q_inv : 0.f;
    for (int w = 0; w < n_warps; ++w) s += s_red[w];
    s_const = s;
  }
  // table: consecutive threads take consecutive centroids of one subspace (coalesced codebook reads)
  for (int e = threadIdx.x; e < M * 256; e += blockDim.x) {
    const int m = e >> 8;
    const float* c = cen + (size_t)e * dsub;
    const float* qq = sqr + m * dsub;
    float a = 0.f;
    for (int i = 0; i < dsub; ++i) {
      const float x = __ldg(c + i);
      if (metric == 0) {
        const float df = qq[i] - x;
        a = fmaf(df, df, a);
      } else {
        a = fmaf(qq[i], x, a);
      }
    }
    table[e] = a;
  }
  __syncthreads();
  const float cst = s_const;
  const bool key_n = metric == 1 && (normalise & 1);             // the reference's l2 branch never normalises the keys (:159-165)
  for (int j0 = warp * KS_ROWS; j0 < k_nn; j0 += n_warps * KS_ROWS) {
    const uint8_t* row[KS_ROWS];
#pragma unroll
    for (int r = 0; r < KS_ROWS; ++r) {
      int64_t id = j0 + r < k_nn ? __ldg(ids + t * k_nn + j0 + r) : 0;
      if (id < 0) id += n_datastore;
      if (id < 0 || id >= n_datastore) id = 0;
      row[r] = codes + id * (int64_t)M;
    }
    float acc[KS_ROWS], nrm[KS_ROWS];
#pragma unroll
    for (int r = 0; r < KS_ROWS; ++r) acc[r] = nrm[r] = 0.f;
    for (int m0 = lane * 4; m0 < M; m0 += 128) {                 // M % 4 == 0: 4 code bytes per lane per pass
      uint32_t cw[KS_ROWS];
#pragma unroll
      for (int r = 0; r < KS_ROWS; ++r) cw[r] = __ldg(reinterpret_cast<const uint32_t*>(row[r] + m0));
#pragma unroll
      for (int r = 0; r < KS_ROWS; ++r) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = ((m0 + i) << 8) + ((cw[r] >> (8 * i)) & 0xff);
          acc[r] += table[e];
          if (key_n) nrm[r] += __ldg(key_norm2 + e);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < KS_ROWS; ++r) {
      const float a = warp_sum(acc[r]);
      const float n2 = key_n ? warp_sum(nrm[r]) : 1.f;
      if (lane == 0 && j0 + r < k_nn) sims[t * k_nn + j0 + r] = metric == 0 ? -(a + cst) : (a + cst) * (key_n ? rsqrtf(n2) : 1.f);
    }
  }
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_knn_sims_keys(const float* queries, int64_t ldq, const void* keys, int32_t key_dtype,
                                       int64_t n_datastore, int32_t d, const int64_t* ids, int64_t k_nn, int32_t metric,
                                       int32_t normalise, float* sims, int64_t T, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(queries && keys && ids && sims, GNNLM_E_ARG, "gnnlm_knn_sims_keys: null pointer");
  GNNLM_CHECK_ARG(key_dtype == GNNLM_F32 || key_dtype == GNNLM_F16, GNNLM_E_UNSUPPORTED, "gnnlm_knn_sims_keys: keys must be fp32 or fp16");
  GNNLM_CHECK_ARG(metric == 0 || metric == 1, GNNLM_E_ARG, "gnnlm_knn_sims_keys: metric must be 0 (l2) or 1 (ip)");
  GNNLM_CHECK_ARG(d > 0 && d % 8 == 0 && d <= 8192 && (uintptr_t)keys % 16 == 0, GNNLM_E_SHAPE,
                  "gnnlm_knn_sims_keys: d must be a multiple of 8 (<= 8192) and keys 16 B aligned");
  GNNLM_CHECK_ARG(n_datastore > 0 && k_nn > 0 && T >= 0 && ldq >= d, GNNLM_E_SHAPE, "gnnlm_knn_sims_keys: bad sizes");
  if (T == 0) return 0;
  const size_t smem = ((size_t)d + KS_THREADS / 32) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (key_dtype == GNNLM_F16)
    knn_sims_keys_kernel<__half><<<(unsigned)T, KS_THREADS, smem, st>>>(queries, ldq, (const __half*)keys, n_datastore, d, ids,
                                                                       (int)k_nn, metric, normalise, sims);
  else
    knn_sims_keys_kernel<float><<<(unsigned)T, KS_THREADS, smem, st>>>(queries, ldq, (const float*)keys, n_datastore, d, ids,
                                                                      (int)k_nn, metric, normalise, sims);
  GNNLM_LAUNCH_CHECK("gnnlm_knn_sims_keys");
  return 0;
}

extern "C" int32_t gnnlm_knn_sims_pq(const float* queries, int64_t ldq, int32_t d_q, const float* rotated, int64_t ldr,
                                     const uint8_t* codes, int64_t n_datastore, int32_t M, int32_t dsub,
                                     const float* centroids, const float* bias, const int64_t* ids, int64_t k_nn,
                                     int32_t metric, const float* key_norm2, int32_t normalise, float* sims, int64_t T,
                                     gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(queries && rotated && codes && centroids && ids && sims, GNNLM_E_ARG, "gnnlm_knn_sims_pq: null pointer");
  GNNLM_CHECK_ARG(metric == 0 || metric == 1, GNNLM_E_ARG, "gnnlm_knn_sims_pq: metric must be 0 (l2) or 1 (ip)");
  GNNLM_CHECK_ARG(normalise >= 0 && normalise <= 3 && (!(normalise & 1) || metric == 0 || key_norm2), GNNLM_E_ARG,
                  "gnnlm_knn_sims_pq: normalise bit 0 (cosine keys) needs key_norm2 [M, 256]");
  GNNLM_CHECK_ARG(M > 0 && M % 4 == 0 && dsub > 0 && n_datastore > 0 && k_nn > 0 && T >= 0, GNNLM_E_SHAPE,
                  "gnnlm_knn_sims_pq: M must be a multiple of 4");
  const size_t smem = ((size_t)M * 256 + (size_t)M * dsub) * sizeof(float);
  GNNLM_CHECK_ARG(smem <= 220 * 1024, GNNLM_E_UNSUPPORTED, "gnnlm_knn_sims_pq: the [M, 256] table (%zu B) does not fit in shared memory", smem);
  if (T == 0) return 0;
  static size_t configured = 0;
  if (smem > configured) {
    GNNLM_CUDA(cudaFuncSetAttribute(knn_sims_pq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  knn_sims_pq_kernel<<<(unsigned)T, 1024, smem, (cudaStream_t)stream>>>(queries, ldq, rotated, ldr, d_q, codes, n_datastore, M,
                                                                       dsub, centroids, bias, ids, (int)k_nn, metric, key_norm2, normalise,
                                                                       sims);
  GNNLM_LAUNCH_CHECK("gnnlm_knn_sims_pq");
  return 0;
}
