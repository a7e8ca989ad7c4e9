// Adaptive-softmax bookkeeping, log-sum-exp finish, kNN-LM interpolation and NLL accumulation.
//
// Replaces (reference paths): AdaptiveSoftmax.adapt_target / the log-softmax + tail-prior adds of
// get_log_prob (fairseq/modules/adaptive_softmax.py:122-145,185-203), KNNModel.get_knn_prob minus
// the faiss search (knn/knn_model.py:187-217), combine_knn_and_vocab_probs
// (fairseq/sequence_scorer.py:55-68,121,135-136), combinetow_probs
// (fairseq/models/transformer.py:1055-1062) and `score_sum += pos_scores.sum()`
// (fairseq_cli/eval_lm.py:273-274).
//
// The GEMM epilogue (gemm_*.cu, *_lse) leaves per-row, per-column-tile (max, sum-exp) partials and
// the picked logit; gnnlm_lse_finish folds them into log-softmax(picked).  gnnlm_knn_mix_nll then
// does, with one warp per token: neighbour softmax over sims/T (online, lanes strided over k_nn),
// the vote on the target (random 4 B gathers from the vals table), the log-space mix and the
// fp64 NLL accumulation -- 12 B read per retrieved neighbour + one 32 B sector per vals gather,
// 4 B written per token.
#include "common.cuh"

namespace gnnlm {

constexpr int MAX_CUT = 8;
struct Cutoffs {
  int64_t c[MAX_CUT];
  int n;
};

// block 0: head_pick for all tokens; block 1+i: ordered row list of tail cluster i
__global__ void __launch_bounds__(1024) adapt_target_kernel(const int64_t* __restrict__ target, int64_t T, Cutoffs cut,
                                                            int32_t* __restrict__ head_pick, int32_t* __restrict__ tail_rows,
                                                            int32_t* __restrict__ tail_pick, int32_t* __restrict__ tail_count) {
  if (blockIdx.x == 0) {
    for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
      const int64_t y = __ldg(target + t);
      int32_t p = (int32_t)y;
      for (int i = 0; i + 1 < cut.n; ++i)
        if (y >= cut.c[i] && y < cut.c[i + 1]) p = (int32_t)cut.c[0] + i;
      if (y < 0 || y >= cut.c[cut.n - 1]) p = -1;
      head_pick[t] = p;
    }
    return;
  }
  const int i = blockIdx.x - 1;
  const int64_t lo = cut.c[i], hi = cut.c[i + 1];
  __shared__ int warp_cnt[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < T; base += blockDim.x) {
    const int64_t t = base + threadIdx.x;
    const int64_t y = t < T ? __ldg(target + t) : -1;
    const bool in = (t < T) && y >= lo && y < hi;
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const int c = warp_cnt[w];
      if (w < wid) before += c;
      total += c;
    }
    const int carry = carry_s;
    if (in) {
      const int pos = carry + before + __popc(bal & ((1u << lane) - 1u));
      tail_rows[(int64_t)i * T + pos] = (int32_t)t;
      tail_pick[(int64_t)i * T + pos] = (int32_t)(y - lo);
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) tail_count[i] = carry_s;
}

// one warp per row: logsumexp over tile partials
__global__ void __launch_bounds__(256) lse_finish_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                                         const float* __restrict__ picked, int64_t n_tiles,
                                                         const int32_t* __restrict__ row_map, float* __restrict__ out,
                                                         int accumulate, int64_t M_cap, const int32_t* __restrict__ m_dev) {
  const int64_t M = live_rows(M_cap, m_dev);
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    float mx = -INFINITY;
    for (int64_t j = lane; j < n_tiles; j += 32) mx = fmaxf(mx, __ldg(part_max + m * n_tiles + j));
    mx = warp_max(mx);
    float s = 0.f;
    for (int64_t j = lane; j < n_tiles; j += 32)
      s += __ldg(part_sum + m * n_tiles + j) * expf(__ldg(part_max + m * n_tiles + j) - mx);
    s = warp_sum(s);
    if (lane == 0) {
      const float lp = __ldg(picked + m) - (mx + logf(s));
      const int64_t r = row_map ? (int64_t)__ldg(row_map + m) : m;
      out[r] = accumulate ? out[r] + lp : lp;
    }
  }
}

__device__ __forceinline__ float logaddexp2(float a, float b) {
  const float mx = fmaxf(a, b);
  if (mx == -INFINITY) return -INFINITY;
  return mx + logf(expf(a - mx) + expf(b - mx));
}

__device__ __forceinline__ int64_t load_val(const void* vals, int val_bytes, int64_t i) {
  return val_bytes == 2 ? (int64_t)__ldg(reinterpret_cast<const int16_t*>(vals) + i)
                        : (int64_t)__ldg(reinterpret_cast<const int32_t*>(vals) + i);
}

__global__ void __launch_bounds__(256) knn_mix_nll_kernel(const float* __restrict__ lm_lp, const float* __restrict__ orig_lp,
                                                          float log_a, float log_1ma, const float* __restrict__ dists,
                                                          const int64_t* __restrict__ ids, int64_t k_nn,
                                                          const void* __restrict__ vals, int val_bytes, int64_t n_datastore,
                                                          const int64_t* __restrict__ target, float sim_sign, float inv_temp,
                                                          float log_lambda, float log_1mlambda, int use_knn,
                                                          const float* __restrict__ weight, int64_t pad_id,
                                                          const int32_t* __restrict__ loss_start, int64_t L,
                                                          float* __restrict__ out_lp,
                                                          float* __restrict__ out_knn_prob, int32_t* __restrict__ out_recall,
                                                          double* __restrict__ nll_acc, int64_t T) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double my_sum = 0.0, my_cnt = 0.0;
  for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < T; t += warps) {
    float lp = __ldg(lm_lp + t);
    if (orig_lp) lp = logaddexp2(__ldg(orig_lp + t) + log_a, lp + log_1ma);         // transformer.py:1055-1062
    if (use_knn) {
      const int64_t y = __ldg(target + t);
      float m = -INFINITY, l = 0.f, h = 0.f;
      int rec = 0;
      for (int64_t j = lane; j < k_nn; j += 32) {
        int64_t id = __ldg(ids + t * k_nn + j);
        float s = __ldg(dists + t * k_nn + j) * sim_sign;                              // knn_model.py:153-157
        if (id == -1) { s = -1e10f; id += n_datastore; }                               // :193 mask; :198 numpy wrap
        if (id < 0 || id >= n_datastore) { s = -1e10f; id = n_datastore - 1; }         // foreign ids: masked, never dereferenced
        s *= inv_temp;                                                                 // :196
        const bool hit = load_val(vals, val_bytes, id) == y;                           // :212
        rec += hit;
        const float mx = fmaxf(m, s);
        const float corr = expf(m - mx), w = expf(s - mx);
        l = l * corr + w;
        h = h * corr + (hit ? w : 0.f);
        m = mx;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
        const float h2 = __shfl_xor_sync(0xffffffffu, h, o);
        rec += __shfl_xor_sync(0xffffffffu, rec, o);
        const float mx = fmaxf(m, m2);
        const float c1 = (m == -INFINITY) ? 0.f : expf(m - mx), c2 = (m2 == -INFINITY) ? 0.f : expf(m2 - mx);
        l = l * c1 + l2 * c2;
        h = h * c1 + h2 * c2;
        m = mx;
      }
      const float p = l > 0.f ? h / l : 0.f;
      if (lane == 0) {
        if (out_knn_prob) out_knn_prob[t] = p;
        if (out_recall) out_recall[t] = rec;
      }
      lp = logaddexp2(lp + log_1mlambda, logf(p + 1e-10f) + log_lambda);               // sequence_scorer.py:55-68,121
    }
    if (lane == 0) {
      if (out_lp) out_lp[t] = lp;
      float w = weight ? __ldg(weight + t) : 1.f;
      if (pad_id >= 0 && target && __ldg(target + t) == pad_id) w = 0.f;               // strip_pad (sequence_scorer.py:159,180)
      if (loss_start && L > 0 && (t % L) < (int64_t)__ldg(loss_start + t / L)) w = 0.f;   // start_indices (:156-162)
      my_sum += (double)lp * (double)w;
      my_cnt += (double)w;
    }
  }
  if (nll_acc) {
    __shared__ double ssum[8], scnt[8];
    if (lane == 0) { ssum[wid] = my_sum; scnt[wid] = my_cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += ssum[w]; b += scnt[w]; }
      if (b != 0.0 || a != 0.0) {
        atomicAdd(nll_acc, a);
        atomicAdd(nll_acc + 1, b);
      }
    }
  }
}

__global__ void __launch_bounds__(256) knn_full_kernel(const float* __restrict__ dists, const int64_t* __restrict__ ids,
                                                       int64_t k_nn, const void* __restrict__ vals, int val_bytes,
                                                       int64_t n_datastore, float sim_sign, float inv_temp,
                                                       float* __restrict__ probs, int64_t V, int64_t T) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < T; t += warps) {
    float m = -INFINITY;
    for (int64_t j = lane; j < k_nn; j += 32) {
      float s = __ldg(dists + t * k_nn + j) * sim_sign;
      { const int64_t id = __ldg(ids + t * k_nn + j); if (id < 0 || id >= n_datastore) s = -1e10f; }
      m = fmaxf(m, s * inv_temp);
    }
    m = warp_max(m);
    float l = 0.f;
    for (int64_t j = lane; j < k_nn; j += 32) {
      float s = __ldg(dists + t * k_nn + j) * sim_sign;
      { const int64_t id = __ldg(ids + t * k_nn + j); if (id < 0 || id >= n_datastore) s = -1e10f; }
      l += expf(s * inv_temp - m);
    }
    l = warp_sum(l);
    for (int64_t j = lane; j < k_nn; j += 32) {
      int64_t id = __ldg(ids + t * k_nn + j);
      float s = __ldg(dists + t * k_nn + j) * sim_sign;
      if (id == -1) { s = -1e10f; id += n_datastore; }
      if (id < 0 || id >= n_datastore) continue;                                      // foreign ids are never dereferenced
      const int64_t val = load_val(vals, val_bytes, id);
      if (val >= 0 && val < V) atomicAdd(probs + t * V + val, expf(s * inv_temp - m) / l);   // knn_model.py:205
    }
  }
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_adapt_target(const int64_t* target, int64_t T, const int64_t* cutoff_host, int32_t n_cut,
                                      int32_t* head_pick, int32_t* tail_rows, int32_t* tail_pick, int32_t* tail_count,
                                      gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(target && cutoff_host && head_pick, GNNLM_E_ARG, "gnnlm_adapt_target: null pointer");
  GNNLM_CHECK_ARG(n_cut >= 1 && n_cut <= MAX_CUT, GNNLM_E_SHAPE, "gnnlm_adapt_target: n_cut must be in [1, %d]", MAX_CUT);
  GNNLM_CHECK_ARG(n_cut == 1 || (tail_rows && tail_pick && tail_count), GNNLM_E_ARG, "gnnlm_adapt_target: tail outputs null");
  if (T == 0) {
    if (n_cut > 1) GNNLM_CUDA(cudaMemsetAsync(tail_count, 0, sizeof(int32_t) * (n_cut - 1), (cudaStream_t)stream));
    return 0;
  }
  Cutoffs c;
  c.n = n_cut;
  for (int i = 0; i < n_cut; ++i) {
    c.c[i] = cutoff_host[i];
    GNNLM_CHECK_ARG(i == 0 || c.c[i] > c.c[i - 1], GNNLM_E_SHAPE, "gnnlm_adapt_target: cutoffs must increase");
  }
  adapt_target_kernel<<<n_cut, 1024, 0, (cudaStream_t)stream>>>(target, T, c, head_pick, tail_rows, tail_pick, tail_count);
  GNNLM_LAUNCH_CHECK("gnnlm_adapt_target");
  return 0;
}

extern "C" int32_t gnnlm_lse_finish(const float* part_max, const float* part_sum, const float* picked, int64_t n_tiles,
                                    const int32_t* row_map, float* out, int32_t accumulate, int64_t M,
                                    const int32_t* m_dev, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(part_max && part_sum && picked && out, GNNLM_E_ARG, "gnnlm_lse_finish: null pointer");
  GNNLM_CHECK_ARG(n_tiles > 0, GNNLM_E_SHAPE, "gnnlm_lse_finish: n_tiles");
  if (M == 0) return 0;
  int64_t blocks = ceil_div(M, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  lse_finish_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(part_max, part_sum, picked, n_tiles, row_map, out,
                                                                         accumulate, M, m_dev);
  GNNLM_LAUNCH_CHECK("gnnlm_lse_finish");
  return 0;
}

extern "C" int32_t gnnlm_knn_mix_nll(const float* lm_lp, const float* orig_lp, float orig_ratio, const float* dists,
                                     const int64_t* ids, int64_t k_nn, const void* vals, int32_t val_bytes,
                                     int64_t n_datastore, const int64_t* target, float sim_sign, float temperature,
                                     float lambda, const float* weight, int64_t pad_id, const int32_t* loss_start,
                                     int64_t L, float* out_lp, float* out_knn_prob, int32_t* out_recall,
                                     double* nll_acc, int64_t T, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(lm_lp, GNNLM_E_ARG, "gnnlm_knn_mix_nll: lm_lp null");
  const int use_knn = (dists != nullptr) && lambda > 0.f;
  if (use_knn) {
    GNNLM_CHECK_ARG(ids && vals && target && k_nn > 0 && (val_bytes == 2 || val_bytes == 4) && n_datastore > 0, GNNLM_E_ARG,
                    "gnnlm_knn_mix_nll: kNN inputs incomplete");
    GNNLM_CHECK_ARG(lambda < 1.f && temperature > 0.f, GNNLM_E_ARG, "gnnlm_knn_mix_nll: need 0 < lambda < 1, T > 0");
  }
  GNNLM_CHECK_ARG(!orig_lp || (orig_ratio > 0.f && orig_ratio < 1.f), GNNLM_E_ARG, "gnnlm_knn_mix_nll: orig_ratio in (0,1)");
  GNNLM_CHECK_ARG((pad_id < 0 && !loss_start) || target, GNNLM_E_ARG, "gnnlm_knn_mix_nll: pad_id / loss_start need target");
  GNNLM_CHECK_ARG(!loss_start || L > 0, GNNLM_E_ARG, "gnnlm_knn_mix_nll: loss_start needs L");
  if (T == 0) return 0;
  int64_t blocks = ceil_div(T, 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  // the reference builds the coefficients with np.log in fp64 and stores them in an fp32 tensor
  const float log_a = orig_lp ? (float)log((double)orig_ratio) : 0.f, log_1ma = orig_lp ? (float)log(1.0 - (double)orig_ratio) : 0.f;
  const float log_l = use_knn ? (float)log((double)lambda) : 0.f, log_1ml = use_knn ? (float)log(1.0 - (double)lambda) : 0.f;
  knn_mix_nll_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      lm_lp, orig_lp, log_a, log_1ma, dists, ids, k_nn, vals, val_bytes, n_datastore, target, sim_sign,
      use_knn ? 1.f / temperature : 1.f, log_l, log_1ml, use_knn, weight, pad_id, loss_start, L, out_lp, out_knn_prob,
      out_recall, nll_acc, T);
  GNNLM_LAUNCH_CHECK("gnnlm_knn_mix_nll");
  return 0;
}

extern "C" int32_t gnnlm_knn_full_prob(const float* dists, const int64_t* ids, int64_t k_nn, const void* vals,
                                       int32_t val_bytes, int64_t n_datastore, float sim_sign, float temperature,
                                       float* probs, int64_t V, int64_t T, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(dists && ids && vals && probs, GNNLM_E_ARG, "gnnlm_knn_full_prob: null pointer");
  GNNLM_CHECK_ARG(k_nn > 0 && V > 0 && temperature > 0.f && (val_bytes == 2 || val_bytes == 4), GNNLM_E_ARG,
                  "gnnlm_knn_full_prob: bad arguments");
  if (T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GNNLM_CUDA(cudaMemsetAsync(probs, 0, sizeof(float) * (size_t)T * V, st));
  int64_t blocks = ceil_div(T, 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  knn_full_kernel<<<(unsigned)blocks, 256, 0, st>>>(dists, ids, k_nn, vals, val_bytes, n_datastore, sim_sign,
                                                    1.f / temperature, probs, V, T);
  GNNLM_LAUNCH_CHECK("gnnlm_knn_full_prob");
  return 0;
}
