// Backward kernels of the HGT fine-tuning step (`--freeze`: only decoder.hgt_decoder.* is trained,
// fairseq/models/transformer_lm.py:183-186; loss = fairseq/criterions/adaptive_loss.py:31-83).
//
// The reference gets these from autograd over DGL's SDDMM / edge_softmax / SpMM (fairseq/models/hgt.py:350-358,383-386),
// F.layer_norm (:405) and F.cross_entropy.  Here: one warp per destination for the segmented softmax + aggregation backward
// (softmax statistics recomputed from Q / K' instead of stored per edge), one block per row for LayerNorm and the
// softmax-cross-entropy of one adaptive-softmax cluster, plus the small data-movement kernels the GEMM backward needs
// (transpose for dW = dY^T X through gnnlm_linear, column sums for the bias, scatter-add for gathered rows).
// First training path: fp32, correctness before speed (three passes over the in-edges, fp32 atomics on dK' / dV').
#include <stdlib.h>

#include "common.cuh"

namespace gnnlm {

// ---------------------------------------------------------------------------------------------- dropout masks
// hgt.py's `drop` (on the output projection, :401) and `attn_drop` (on the edge-softmax weights, one draw per edge and head, :356)
// and the adaptive softmax's input / tail dropouts draw from torch's generator in the reference; here a mask is a pure function
// of (seed, element) -- splitmix64 of seed + index * golden ratio, top 24 bits against p -- so that forward and backward
// regenerate it instead of storing it, and the tests can replay the same mask through the oracle.
__host__ __device__ __forceinline__ uint64_t dm_mix(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// multiplier of a kept element: 1 / (1 - p); dropped: 0.  p_thresh = floor(p * 2^24)
__device__ __forceinline__ float dm_scale(uint64_t seed, uint64_t idx, uint32_t p_thresh, float keep_scale) {
  return (uint32_t)(dm_mix(seed, idx) >> 40) >= p_thresh ? keep_scale : 0.f;
}
__device__ __forceinline__ uint64_t dm_edge(int64_t dst, int64_t src, int head) {
  return ((uint64_t)dst << 38) ^ ((uint64_t)src << 6) ^ (uint64_t)head;
}
struct AttnDrop {
  uint64_t seed;
  uint32_t p_thresh;      // 0: no dropout
  float keep_scale;
};

// ---------------------------------------------------------------------------------------------- edge attention backward
// forward (edge_attn.cu):  out[v] = scale * sum_e alpha_e V'[u_e],  alpha = softmax_e <Q[v,h], K'[u_e,h]>   per head h
// backward, per destination v and head h, with g_e = <dout[v,h], V'[u_e,h]> and D = sum_e alpha_e g_e:
//   dV'[u_e] += scale * alpha_e dout[v]     ds_e = scale * alpha_e (g_e - D)     dQ[v] = sum_e ds_e K'[u_e]     dK'[u_e] += ds_e Q[v]
// Edge sources: CSR (indptr / indices; indices == NULL: source = edge id) or, causal_L > 0, the implicit causal range of
// token_block_dataset.py:586-594 inside blocks of causal_L destinations.
template <int C>
__global__ void __launch_bounds__(256) edge_attn_bwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                            int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                            const float* __restrict__ dout, int64_t ldo,
                                                            const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                            const int32_t* __restrict__ dst_ids, int64_t n_dst_cap,
                                                            const int32_t* __restrict__ n_dst_dev, int64_t causal_L, int64_t intra_ctx,
                                                            int group, float scale, float* __restrict__ dq, int64_t lddq,
                                                            float* __restrict__ dk, int64_t lddk, float* __restrict__ dv, int64_t lddv,
                                                            AttnDrop ad, float* __restrict__ stats, int H, int exclusive) {
  // stats != NULL: first pass of the atomic-free form -- dQ and {max, 1 / sum, D} per (destination, head) only; dK' / dV' come
  // from csr_bwd_dkv_kernel.  exclusive: every source feeds exactly one edge (source id = edge id): plain stores.
  const int64_t n_dst = live_rows(n_dst_cap, n_dst_dev);
  const int lane = threadIdx.x & 31;
  const int col = lane * C;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_dst; i += warps) {
    int64_t e0, e1;
    if (causal_L > 0) {
      const int64_t b0 = (i / causal_L) * causal_L;
      e0 = intra_ctx > 0 && i - intra_ctx + 1 > b0 ? i - intra_ctx + 1 : b0;
      e1 = i + 1;
    } else {
      const int64_t row = dst_ids ? (int64_t)__ldg(dst_ids + i) : i;
      e0 = __ldg(indptr + row);
      e1 = __ldg(indptr + row + 1);
    }
    float qr[C], gr[C], dqr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      qr[c] = q[i * ldq + col + c];
      gr[c] = dout[i * ldo + col + c];
      dqr[c] = 0.f;
    }
    auto src_of = [&](int64_t e) { return causal_L > 0 ? e : (indices ? (int64_t)__ldg(indices + e) : e); };
    auto head_dot = [&](const float (&a)[C], const float* __restrict__ row) {
      float p = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) p = fmaf(a[c], row[col + c], p);
      for (int o = group >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      return p;
    };
    // pass A: softmax statistics
    float m = -INFINITY, l = 0.f;
    for (int64_t e = e0; e < e1; ++e) {
      const float s = head_dot(qr, k + src_of(e) * ldk);
      const float mx = fmaxf(m, s);
      l = l * __expf(m - mx) + __expf(s - mx);
      m = mx;
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    // attention dropout (hgt.py:356): out = sum_e alpha_e beta_e V'_e with beta_e in {0, 1 / (1 - p)}; then g_e -> beta_e g_e
    const int head = lane / group;
    auto beta = [&](int64_t u) { return ad.p_thresh ? dm_scale(ad.seed, dm_edge(i, u, head), ad.p_thresh, ad.keep_scale) : 1.f; };
    // pass B: D = sum_e alpha_e beta_e g_e
    float D = 0.f;
    for (int64_t e = e0; e < e1; ++e) {
      const int64_t u = src_of(e);
      const float a = __expf(head_dot(qr, k + u * ldk) - m) * inv_l;
      D = fmaf(a * beta(u), head_dot(gr, v + u * ldv), D);
    }
    // pass C: gradients
    for (int64_t e = e0; e < e1; ++e) {
      const int64_t u = src_of(e);
      const float* kr = k + u * ldk;
      const float a = __expf(head_dot(qr, kr) - m) * inv_l;
      const float be = beta(u);
      const float ds = scale * a * (be * head_dot(gr, v + u * ldv) - D);
      const float av = scale * a * be;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dqr[c] = fmaf(ds, kr[col + c], dqr[c]);
        if (stats) continue;
        if (exclusive) {
          dk[u * lddk + col + c] = ds * qr[c];
          dv[u * lddv + col + c] = av * gr[c];
        } else {
          atomicAdd(dk + u * lddk + col + c, ds * qr[c]);
          atomicAdd(dv + u * lddv + col + c, av * gr[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dq[i * lddq + col + c] = dqr[c];
    if (stats && (lane % group) == 0) {
      float* st = stats + (i * H + head) * 3;
      st[0] = m; st[1] = inv_l; st[2] = D;
    }
  }
}

// Second pass of the atomic-free CSR form for SYMMETRIC edge sets (u -> v exists iff v -> u does, with the same multiplicity --
// build_ntgt_edges(bidirect=True) + self loops, token_block_dataset.py:395-400): the in-edge list of a node is then its out-edge
// list as well, so one warp per SOURCE u walks row u of the same CSR, reads the statistics of each destination it feeds and
// accumulates dK'[u] / dV'[u] in registers.
template <int C>
__global__ void __launch_bounds__(256) csr_bwd_dkv_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                                                          const float* __restrict__ v, int64_t ldv, const float* __restrict__ dout,
                                                          int64_t ldo, const int32_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ indices, int64_t n, int group, int H, float scale,
                                                          const float* __restrict__ stats, float* __restrict__ dk, int64_t lddk,
                                                          float* __restrict__ dv, int64_t lddv, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int col = lane * C;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n; u += warps) {
    float kr[C], vr[C], dkr[C], dvr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      kr[c] = k[u * ldk + col + c];
      vr[c] = v[u * ldv + col + c];
      dkr[c] = dvr[c] = 0.f;
    }
    const int head = lane / group;
    const int64_t e1 = __ldg(indptr + u + 1);
    for (int64_t e = __ldg(indptr + u); e < e1; ++e) {
      const int64_t i = __ldg(indices + e);                       // a destination u feeds
      const float* qr = q + i * ldq;
      const float* gr = dout + i * ldo;
      float sc = 0.f, g = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        sc = fmaf(kr[c], qr[col + c], sc);
        g = fmaf(vr[c], gr[col + c], g);
      }
      for (int o = group >> 1; o > 0; o >>= 1) {
        sc += __shfl_xor_sync(0xffffffffu, sc, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
      }
      const float* st = stats + (i * H + head) * 3;
      const float a = scale * __expf(sc - st[0]) * st[1];
      const float be = ad.p_thresh ? dm_scale(ad.seed, dm_edge(i, u, head), ad.p_thresh, ad.keep_scale) : 1.f;
      const float ds = a * (be * g - st[2]);
      const float ab = a * be;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dkr[c] = fmaf(ds, qr[col + c], dkr[c]);
        dvr[c] = fmaf(ab, gr[col + c], dvr[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      dk[u * lddk + col + c] = dkr[c];
      dv[u * lddv + col + c] = dvr[c];
    }
  }
}

// Implicit causal edges (4.7 M per 3072-token block and layer) without atomics: a by-destination pass produces dQ and the per
// (destination, head) softmax statistics {max, 1 / sum, D}; a by-source pass recomputes alpha from them and accumulates dK' / dV'
// of its source row in registers.  (The CSR kernel above on the same edges spends 0.2 s per Wiki103 block in fp32 atomics.)
template <int C>
__global__ void __launch_bounds__(256) causal_bwd_dq_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                            int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                            const float* __restrict__ dout, int64_t ldo, int64_t T, int64_t Lb,
                                                            int64_t intra_ctx, int group, int H, float scale, float* __restrict__ dq,
                                                            int64_t lddq, float* __restrict__ stats, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int col = lane * C;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // heaviest destinations (end of a block) first
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < T; it += warps) {
    const int64_t i = T - 1 - it;
    const int64_t b0 = (i / Lb) * Lb;
    const int64_t e0 = intra_ctx > 0 && i - intra_ctx + 1 > b0 ? i - intra_ctx + 1 : b0;
    float qr[C], gr[C], dqr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      qr[c] = q[i * ldq + col + c];
      gr[c] = dout[i * ldo + col + c];
      dqr[c] = 0.f;
    }
    auto head_dot = [&](const float (&a)[C], const float* __restrict__ row) {
      float p = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) p = fmaf(a[c], row[col + c], p);
      for (int o = group >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      return p;
    };
    const int head = lane / group;
    auto beta = [&](int64_t u) { return ad.p_thresh ? dm_scale(ad.seed, dm_edge(i, u, head), ad.p_thresh, ad.keep_scale) : 1.f; };
    float m = -INFINITY, l = 0.f, dn = 0.f;
    for (int64_t u = e0; u <= i; ++u) {
      const float s = head_dot(qr, k + u * ldk), g = beta(u) * head_dot(gr, v + u * ldv);
      const float mx = fmaxf(m, s), corr = __expf(m - mx), w = __expf(s - mx);
      l = l * corr + w;
      dn = dn * corr + w * g;
      m = mx;
    }
    const float inv_l = 1.f / l, D = dn * inv_l;              // every destination has its self edge: l > 0
    if ((lane % group) == 0) {
      float* st = stats + (i * H + lane / group) * 3;
      st[0] = m; st[1] = inv_l; st[2] = D;
    }
    for (int64_t u = e0; u <= i; ++u) {
      const float* kr = k + u * ldk;
      const float a = __expf(head_dot(qr, kr) - m) * inv_l;
      const float ds = scale * a * (beta(u) * head_dot(gr, v + u * ldv) - D);
#pragma unroll
      for (int c = 0; c < C; ++c) dqr[c] = fmaf(ds, kr[col + c], dqr[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dq[i * lddq + col + c] = dqr[c];
  }
}

template <int C>
__global__ void __launch_bounds__(256) causal_bwd_dkv_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                             int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                             const float* __restrict__ dout, int64_t ldo, int64_t T, int64_t Lb,
                                                             int64_t intra_ctx, int group, int H, float scale,
                                                             const float* __restrict__ stats, float* __restrict__ dk, int64_t lddk,
                                                             float* __restrict__ dv, int64_t lddv, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int col = lane * C;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // heaviest sources (start of a block) first: warp index -> position inside the block, interleaved over the blocks
  const int64_t B = T / Lb;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < T; it += warps) {
    const int64_t u = (it % B) * Lb + it / B;
    const int64_t b1 = (u / Lb + 1) * Lb;
    const int64_t v1 = intra_ctx > 0 && u + intra_ctx < b1 ? u + intra_ctx : b1;      // destinations u <= v < v1
    float kr[C], vr[C], dkr[C], dvr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      kr[c] = k[u * ldk + col + c];
      vr[c] = v[u * ldv + col + c];
      dkr[c] = dvr[c] = 0.f;
    }
    const int head = lane / group;
    for (int64_t i = u; i < v1; ++i) {
      const float* qr = q + i * ldq;
      const float* gr = dout + i * ldo;
      float s = 0.f, g = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        s = fmaf(kr[c], qr[col + c], s);
        g = fmaf(vr[c], gr[col + c], g);
      }
      for (int o = group >> 1; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
      }
      const float* st = stats + (i * H + head) * 3;
      const float a = scale * __expf(s - st[0]) * st[1];
      const float be = ad.p_thresh ? dm_scale(ad.seed, dm_edge(i, u, head), ad.p_thresh, ad.keep_scale) : 1.f;
      const float ds = a * (be * g - st[2]);
      const float ab = a * be;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dkr[c] = fmaf(ds, qr[col + c], dkr[c]);
        dvr[c] = fmaf(ab, gr[col + c], dvr[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      dk[u * lddk + col + c] = dkr[c];
      dv[u * lddv + col + c] = dvr[c];
    }
  }
}

// Training forward with attention dropout: out[v] (+)= scale * sum_e alpha_e beta_e V'[u_e] (two passes over the in-edges; the
// evaluation kernels of edge_attn.cu have no dropout).  Same edge sources as edge_attn_bwd_kernel.
template <int C>
__global__ void __launch_bounds__(256) edge_attn_train_fwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                                  int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                                  const int32_t* __restrict__ indptr,
                                                                  const int32_t* __restrict__ indices, int64_t n_dst, int64_t causal_L,
                                                                  int64_t intra_ctx, int group, float scale, int accumulate,
                                                                  float* __restrict__ out, int64_t ldo, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int col = lane * C;
  const int head = lane / group;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_dst; it += warps) {
    const int64_t i = causal_L > 0 ? n_dst - 1 - it : it;
    int64_t e0, e1;
    if (causal_L > 0) {
      const int64_t b0 = (i / causal_L) * causal_L;
      e0 = intra_ctx > 0 && i - intra_ctx + 1 > b0 ? i - intra_ctx + 1 : b0;
      e1 = i + 1;
    } else {
      e0 = __ldg(indptr + i);
      e1 = __ldg(indptr + i + 1);
    }
    float qr[C], acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      qr[c] = q[i * ldq + col + c];
      acc[c] = 0.f;
    }
    auto src_of = [&](int64_t e) { return causal_L > 0 ? e : (indices ? (int64_t)__ldg(indices + e) : e); };
    auto score = [&](int64_t u) {
      const float* row = k + u * ldk;
      float p = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) p = fmaf(qr[c], row[col + c], p);
      for (int o = group >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      return p;
    };
    float m = -INFINITY, l = 0.f;
    for (int64_t e = e0; e < e1; ++e) {
      const float s = score(src_of(e));
      const float mx = fmaxf(m, s);
      l = l * __expf(m - mx) + __expf(s - mx);
      m = mx;
    }
    const float inv_l = l > 0.f ? scale / l : 0.f;
    for (int64_t e = e0; e < e1; ++e) {
      const int64_t u = src_of(e);
      const float w = __expf(score(u) - m) * inv_l * (ad.p_thresh ? dm_scale(ad.seed, dm_edge(i, u, head), ad.p_thresh, ad.keep_scale) : 1.f);
      const float* vr = v + u * ldv;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(w, vr[col + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[i * ldo + col + c] = (accumulate ? out[i * ldo + col + c] : 0.f) + acc[c];
  }
}

// y = x * mask / (1 - p), mask = f(seed, row * cols + col): the forward of nn.Dropout in training mode and, applied to dy, its backward
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy,
                                                      int64_t rows, int64_t cols, uint64_t seed, uint32_t p_thresh, float keep_scale) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / cols, c = i % cols;
    y[r * ldy + c] = x[r * ldx + c] * dm_scale(seed, (uint64_t)i, p_thresh, keep_scale);
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm backward
// y = LayerNorm(o + res) * gamma + beta (hgt.py:403-405).  dx = d(o) = d(res); dgamma / dbeta accumulated with atomics.
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ o, int64_t ldo_, const float* __restrict__ res,
                                                            int64_t ldr, const float* __restrict__ gamma, float eps,
                                                            const float* __restrict__ dy, int64_t ldy, int64_t rows_cap,
                                                            const int32_t* __restrict__ rows_dev, int d, float* __restrict__ dx,
                                                            int64_t ldx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int MAXC = 8;                                    // d <= 2048
  __shared__ float red[8];
  const int64_t rows = live_rows(rows_cap, rows_dev);
  float ag[MAXC], ab[MAXC];
#pragma unroll
  for (int j = 0; j < MAXC; ++j) ag[j] = ab[j] = 0.f;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    float x[MAXC], g[MAXC];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const int c = threadIdx.x + j * 256;
      x[j] = c < d ? o[r * ldo_ + c] + (res ? res[r * ldr + c] : 0.f) : 0.f;
      s += x[j];
    }
    const float mean = block_sum_256(s, red) / d;
    float vs = 0.f;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const int c = threadIdx.x + j * 256;
      x[j] = c < d ? x[j] - mean : 0.f;
      vs += x[j] * x[j];
    }
    const float rstd = rsqrtf(block_sum_256(vs, red) / d + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const int c = threadIdx.x + j * 256;
      x[j] *= rstd;                                          // x hat
      const float dyc = c < d ? dy[r * ldy + c] : 0.f;
      g[j] = c < d ? dyc * gamma[c] : 0.f;
      ag[j] += dyc * x[j];
      ab[j] += dyc;
      s1 += g[j];
      s2 += g[j] * x[j];
    }
    const float m1 = block_sum_256(s1, red) / d, m2 = block_sum_256(s2, red) / d;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const int c = threadIdx.x + j * 256;
      if (c < d) dx[r * ldx + c] = rstd * (g[j] - m1 - x[j] * m2);
    }
  }
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    const int c = threadIdx.x + j * 256;
    if (c < d) {
      atomicAdd(dgamma + c, ag[j]);
      atomicAdd(dbeta + c, ab[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- softmax cross-entropy
// One cluster of the adaptive softmax (adaptive_softmax.py:147-168 + F.cross_entropy(reduction='sum'), adaptive_loss.py:62-70):
// loss += -log softmax(logits[r])[target[r]];  logits[r] <- grad_scale * (softmax(logits[r]) - onehot(target[r])).
// target < 0: the row is ignored (zero gradient).
__global__ void __launch_bounds__(256) xent_fwd_bwd_kernel(float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ target,
                                                           int64_t rows, int64_t C, float grad_scale, double* __restrict__ loss) {
  __shared__ float red[8];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    float* x = logits + r * ld;
    const int64_t t = target[r];
    if (t < 0 || t >= C) {
      for (int64_t c = threadIdx.x; c < C; c += 256) x[c] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int64_t c = threadIdx.x; c < C; c += 256) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    float s = 0.f;
    for (int64_t c = threadIdx.x; c < C; c += 256) s += __expf(x[c] - mx);
    const float l = block_sum_256(s, red);
    const float xt = x[t];
    __syncthreads();
    const float inv = grad_scale / l;
    for (int64_t c = threadIdx.x; c < C; c += 256) x[c] = __expf(x[c] - mx) * inv - (c == t ? grad_scale : 0.f);
    if (threadIdx.x == 0) atomicAdd(loss, (double)(mx + logf(l) - xt));
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- data movement
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows_cap,
                                                        const int32_t* __restrict__ rows_dev, int64_t cols, float* __restrict__ dst,
                                                        int64_t ld_dst, int64_t rows_pad) {
  __shared__ float tile[32][33];
  const int64_t rows = live_rows(rows_cap, rows_dev);
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;     // rows past the live count: zeros (k-padding of dW)
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t c = c0 + j, r = r0 + tx;
    if (c < cols && r < rows_pad) dst[c * ld_dst + r] = tile[tx][j];
  }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t ld, int64_t rows_cap,
                                                     const int32_t* __restrict__ rows_dev, int64_t cols, float* __restrict__ out) {
  const int64_t rows = live_rows(rows_cap, rows_dev);
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) s += x[r * ld + c];
  atomicAdd(out + c, s);
}

__global__ void __launch_bounds__(256) axpy_kernel(float* __restrict__ y, int64_t ldy, const float* __restrict__ x, int64_t ldx,
                                                   int64_t rows_cap, const int32_t* __restrict__ rows_dev, int64_t cols, float a) {
  const int64_t n = live_rows(rows_cap, rows_dev) * cols;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / cols, c = i % cols;
    y[r * ldy + c] += a * x[r * ldx + c];
  }
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(float* __restrict__ dst, int64_t ld_dst, const float* __restrict__ src,
                                                               int64_t ld_src, const int32_t* __restrict__ ids, int64_t rows_cap,
                                                               const int32_t* __restrict__ rows_dev, int64_t cols) {
  const int64_t n = live_rows(rows_cap, rows_dev) * cols;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / cols, c = i % cols;
    atomicAdd(dst + (int64_t)__ldg(ids + r) * ld_dst + c, src[r * ld_src + c]);
  }
}

// ---------------------------------------------------------------------------------------------- causal edges, GEMM form
// The backward of the tgt-intra-tgt attention on the tensor cores at fp32 parity (MATH_F16X3 products through
// gnnlm_linear_batched_f16x3, like the forward of attn_gemm.cu): per block and head
//     S = Q K'^T            g = dOut V'^T                               (two products, causal tile schedule)
//     P = softmax_causal(S)   D_i = sum_j P_ij beta_ij g_ij   w_ij = scale beta_ij P_ij   ds_ij = scale P_ij (beta_ij g_ij - D_i)
//     dQ = ds K'            dV' = w^T dOut            dK' = ds^T Q      (three products)
// and the two kernels below are everything in between: a row pass for the statistics {max, 1 / sum, D} and a 64 x 64 tile pass
// that writes ds row-major and w, ds TRANSPOSED, all as split-fp16 A operands [H, L, 2L].  Only tiles on or below the diagonal are
// touched: the buffers are zero above it once and stay so (the caller keeps them).  beta = the attention dropout multiplier of the
// edge (dm_scale; 1 without dropout), regenerated from (seed, destination, source, head) as in the forward.
__global__ void __launch_bounds__(256) causal_bwd_stats_kernel(const float* __restrict__ S, const float* __restrict__ G, int64_t L,
                                                               int64_t ctx, int H, int64_t row0, float* __restrict__ stats, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)H * L;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const int64_t i = r % L;
    const int head = (int)(r / L);
    const float* s = S + r * L;
    const float* g = G + r * L;
    const int64_t lo_j = (ctx > 0 && i + 1 > ctx) ? i + 1 - ctx : 0;
    float m = -INFINITY, l = 0.f;
    for (int64_t j = (lo_j & ~(int64_t)3) + lane * 4; j <= i; j += 128) {
      const float4 x4 = *reinterpret_cast<const float4*>(s + j);
      float x[4] = {x4.x, x4.y, x4.z, x4.w};
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (j + e < lo_j || j + e > i) x[e] = -INFINITY;
        mx = fmaxf(mx, x[e]);
      }
      if (mx > m) {
        l *= __expf(m - mx);
        m = mx;
      }
      if (m > -INFINITY) {
#pragma unroll
        for (int e = 0; e < 4; ++e) l += __expf(x[e] - m);
      }
    }
    const float M = warp_max(m);                                  // the diagonal is always valid: M is finite
    l = warp_sum(m > -INFINITY ? l * __expf(m - M) : 0.f);
    const float inv = 1.f / l;
    float D = 0.f;
    for (int64_t j = (lo_j & ~(int64_t)3) + lane * 4; j <= i; j += 128) {
      const float4 x4 = *reinterpret_cast<const float4*>(s + j), g4 = *reinterpret_cast<const float4*>(g + j);
      const float x[4] = {x4.x, x4.y, x4.z, x4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (j + e >= lo_j && j + e <= i) {
          const float be = ad.p_thresh ? dm_scale(ad.seed, dm_edge(row0 + i, row0 + j + e, head), ad.p_thresh, ad.keep_scale) : 1.f;
          D = fmaf(__expf(x[e] - M) * inv, be * gg[e], D);
        }
    }
    D = warp_sum(D);
    if (lane == 0) {
      stats[r * 3 + 0] = M;
      stats[r * 3 + 1] = inv;
      stats[r * 3 + 2] = D;
    }
  }
}

constexpr int CBT = 64;                                           // tile edge of the transposing pass
__global__ void __launch_bounds__(256) causal_bwd_tile_kernel(const float* __restrict__ S, const float* __restrict__ G,
                                                              const float* __restrict__ stats, int64_t L, int64_t ctx, int64_t row0,
                                                              float scale, __half* __restrict__ dS, __half* __restrict__ WT,
                                                              __half* __restrict__ dST, AttnDrop ad) {
  const int jt = blockIdx.x, it = blockIdx.y, head = blockIdx.z;
  if (jt > it) return;                                            // above the diagonal: zero, and never written
  __shared__ float sw[CBT][CBT + 1], sd[CBT][CBT + 1];
  const int64_t i0 = (int64_t)it * CBT, j0 = (int64_t)jt * CBT;
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int il = r0 + 16 * rr;
    const int64_t i = i0 + il, row = (int64_t)head * L + i;
    const float M = stats[row * 3], inv = stats[row * 3 + 1], D = stats[row * 3 + 2];
    const int64_t lo_j = (ctx > 0 && i + 1 > ctx) ? i + 1 - ctx : 0;
    const float4 x4 = *reinterpret_cast<const float4*>(S + row * L + j0 + c4), g4 = *reinterpret_cast<const float4*>(G + row * L + j0 + c4);
    const float x[4] = {x4.x, x4.y, x4.z, x4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float w[4], ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t j = j0 + c4 + e;
      w[e] = ds[e] = 0.f;
      if (j >= lo_j && j <= i) {                                  // entries of S / G beyond the diagonal may be unwritten: never used
        const float be = ad.p_thresh ? dm_scale(ad.seed, dm_edge(row0 + i, row0 + j, head), ad.p_thresh, ad.keep_scale) : 1.f;
        const float P = __expf(x[e] - M) * inv;
        w[e] = scale * be * P;
        ds[e] = scale * P * (be * gg[e] - D);
      }
      sw[il][c4 + e] = w[e];
      sd[il][c4 + e] = ds[e];
    }
    uint2 hi, lo;
    split4_f16(ds[0], ds[1], ds[2], ds[3], hi, lo);
    __half* p = dS + row * 2 * L + j0 + c4;
    *reinterpret_cast<uint2*>(p) = hi;
    *reinterpret_cast<uint2*>(p + L) = lo;
  }
  __syncthreads();
  const int iq = (threadIdx.x & 15) * 4;                          // four consecutive destinations of one source row
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int jl = r0 + 16 * rr;
    const int64_t trow = ((int64_t)head * L + j0 + jl) * 2 * L + i0 + iq;
    uint2 hi, lo;
    split4_f16(sw[iq][jl], sw[iq + 1][jl], sw[iq + 2][jl], sw[iq + 3][jl], hi, lo);
    *reinterpret_cast<uint2*>(WT + trow) = hi;
    *reinterpret_cast<uint2*>(WT + trow + L) = lo;
    split4_f16(sd[iq][jl], sd[iq + 1][jl], sd[iq + 2][jl], sd[iq + 3][jl], hi, lo);
    *reinterpret_cast<uint2*>(dST + trow) = hi;
    *reinterpret_cast<uint2*>(dST + trow + L) = lo;
  }
}

// ---------------------------------------------------------------------------------------------- ntgt-intra-ntgt chains
// Backward of gnnlm_hgt_cluster_attn (all-nodes form): every valid (token, neighbour) pair owns a chain of w <= 7 contiguous
// node ids whose node at sorted position p attends to positions p-1, p, p+1 (cluster_attn.cu).  One warp per (cluster, head)
// walks the chain once with a three-row window of K' / V' and their gradient accumulators in registers: every row of Q, K', V',
// dOut is read once and every row of dQ, dK', dV' written once (512 B contiguous per warp and row at d_k = 128), no atomics.
// (The CSR kernel on the same edges: 23 ms per Wiki103 layer, 1.4 G fp32 atomics.)
template <int C>
__global__ void __launch_bounds__(256) cluster_attn_bwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                               int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                               const float* __restrict__ dout, int64_t ldo,
                                                               const int32_t* __restrict__ node_base, const int32_t* __restrict__ cluster_nl,
                                                               int64_t n_clusters, int H, float scale, float* __restrict__ dq, int64_t lddq,
                                                               float* __restrict__ dk, int64_t lddk, float* __restrict__ dv, int64_t lddv,
                                                               AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int64_t n_items = n_clusters * H;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  auto load = [&](const float* __restrict__ p, float (&r)[C]) {
    if constexpr (C == 4) {
      const float4 x = *reinterpret_cast<const float4*>(p);
      r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) r[c] = p[c];
    }
  };
  auto store = [&](float* __restrict__ p, const float (&r)[C]) {
    if constexpr (C == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) p[c] = r[c];
    }
  };
  auto dot = [&](const float (&a)[C], const float (&b)[C]) {
    float p = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) p = fmaf(a[c], b[c], p);
    return warp_sum(p);
  };
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += warps) {
    const int64_t cl = it / H;
    const int head = (int)(it - cl * H);
    const int base = __ldg(node_base + cl);
    const int w = __ldg(node_base + cl + 1) - base;
    if (w <= 0) continue;
    const int nl = __ldg(cluster_nl + cl);
    const int col = head * 32 * C + lane * C;
    auto id_of = [&](int p) { return (int64_t)base + (p == nl ? 0 : (p < nl ? p + 1 : p)); };
    float Kp[C], Vp[C], Kc[C], Vc[C], Kn[C], Vn[C], dKp[C], dVp[C], dKc[C], dVc[C], dKn[C], dVn[C];
#pragma unroll
    for (int c = 0; c < C; ++c) Kp[c] = Vp[c] = Kn[c] = Vn[c] = dKp[c] = dVp[c] = dKc[c] = dVc[c] = dKn[c] = dVn[c] = 0.f;
    int64_t id_p = 0, id_c = id_of(0), id_n = 0;
    load(k + id_c * ldk + col, Kc);
    load(v + id_c * ldv + col, Vc);
    for (int p = 0; p < w; ++p) {
      const bool has_l = p > 0, has_r = p + 1 < w;
      if (has_r) {
        id_n = id_of(p + 1);
        load(k + id_n * ldk + col, Kn);
        load(v + id_n * ldv + col, Vn);
      }
      float qr[C], gr[C];
      load(q + id_c * ldq + col, qr);
      load(dout + id_c * ldo + col, gr);
      const float s1 = dot(qr, Kc), s0 = has_l ? dot(qr, Kp) : -INFINITY, s2 = has_r ? dot(qr, Kn) : -INFINITY;
      const float m = fmaxf(s1, fmaxf(s0, s2));
      const float e0 = has_l ? __expf(s0 - m) : 0.f, e1 = __expf(s1 - m), e2 = has_r ? __expf(s2 - m) : 0.f;
      const float inv = 1.f / (e0 + e1 + e2);
      auto beta = [&](int64_t u) { return ad.p_thresh ? dm_scale(ad.seed, dm_edge(id_c, u, head), ad.p_thresh, ad.keep_scale) : 1.f; };
      const float b0 = has_l ? beta(id_p) : 0.f, b1 = beta(id_c), b2 = has_r ? beta(id_n) : 0.f;
      const float g1 = dot(gr, Vc), g0 = has_l ? dot(gr, Vp) : 0.f, g2 = has_r ? dot(gr, Vn) : 0.f;
      const float a0 = e0 * inv, a1 = e1 * inv, a2 = e2 * inv;
      const float D = a0 * b0 * g0 + a1 * b1 * g1 + a2 * b2 * g2;
      const float ds0 = scale * a0 * (b0 * g0 - D), ds1 = scale * a1 * (b1 * g1 - D), ds2 = scale * a2 * (b2 * g2 - D);
      const float av0 = scale * a0 * b0, av1 = scale * a1 * b1, av2 = scale * a2 * b2;
      float dqr[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dqr[c] = ds0 * Kp[c] + ds1 * Kc[c] + ds2 * Kn[c];
        dKp[c] = fmaf(ds0, qr[c], dKp[c]); dVp[c] = fmaf(av0, gr[c], dVp[c]);
        dKc[c] = fmaf(ds1, qr[c], dKc[c]); dVc[c] = fmaf(av1, gr[c], dVc[c]);
        dKn[c] = fmaf(ds2, qr[c], dKn[c]); dVn[c] = fmaf(av2, gr[c], dVn[c]);
      }
      store(dq + id_c * lddq + col, dqr);
      if (has_l) {                                                 // position p-1 has received from p-2, p-1, p: complete
        store(dk + id_p * lddk + col, dKp);
        store(dv + id_p * lddv + col, dVp);
      }
      id_p = id_c; id_c = id_n;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        Kp[c] = Kc[c]; Vp[c] = Vc[c]; dKp[c] = dKc[c]; dVp[c] = dVc[c];
        Kc[c] = Kn[c]; Vc[c] = Vn[c]; dKc[c] = dKn[c]; dVc[c] = dVn[c];
        Kn[c] = Vn[c] = dKn[c] = dVn[c] = 0.f;
      }
    }
    store(dk + id_p * lddk + col, dKp);                            // the last position
    store(dv + id_p * lddv + col, dVp);
  }
}

// Backward of an edge type whose sources are CONTIGUOUS and EXCLUSIVE per destination (indices == NULL: source id = edge id -- the
// ('ntgt','inter','tgt') edges over the compact centre rows): one warp per (destination, head), 32 * C = d_k features per row-head
// (512 B contiguous at d_k = 128), two passes over the destination's rows (the second from L1 / L2), plain stores for dK' / dV'.
template <int C>
__global__ void __launch_bounds__(256) ranged_attn_bwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                              int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                              const float* __restrict__ dout, int64_t ldo,
                                                              const int32_t* __restrict__ indptr, int64_t n_dst, int H, float scale,
                                                              float* __restrict__ dq, int64_t lddq, float* __restrict__ dk, int64_t lddk,
                                                              float* __restrict__ dv, int64_t lddv, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int64_t n_items = n_dst * H;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  auto load = [&](const float* __restrict__ p, float (&r)[C]) {
    if constexpr (C == 4) {
      const float4 x = *reinterpret_cast<const float4*>(p);
      r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) r[c] = p[c];
    }
  };
  auto store = [&](float* __restrict__ p, const float (&r)[C]) {
    if constexpr (C == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) p[c] = r[c];
    }
  };
  auto dot = [&](const float (&a)[C], const float (&b)[C]) {
    float p = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) p = fmaf(a[c], b[c], p);
    return warp_sum(p);
  };
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += warps) {
    const int64_t i = it / H;
    const int head = (int)(it - i * H);
    const int col = head * 32 * C + lane * C;
    const int64_t e0 = __ldg(indptr + i), e1 = __ldg(indptr + i + 1);
    float qr[C], gr[C], dqr[C];
    load(q + i * ldq + col, qr);
    load(dout + i * ldo + col, gr);
#pragma unroll
    for (int c = 0; c < C; ++c) dqr[c] = 0.f;
    auto beta = [&](int64_t u) { return ad.p_thresh ? dm_scale(ad.seed, dm_edge(i, u, head), ad.p_thresh, ad.keep_scale) : 1.f; };
    float m = -INFINITY, l = 0.f, dn = 0.f;
    for (int64_t u = e0; u < e1; ++u) {
      float kr[C], vr[C];
      load(k + u * ldk + col, kr);
      load(v + u * ldv + col, vr);
      const float sc = dot(qr, kr), g = beta(u) * dot(gr, vr);
      const float mx = fmaxf(m, sc), corr = __expf(m - mx), w = __expf(sc - mx);
      l = l * corr + w;
      dn = dn * corr + w * g;
      m = mx;
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f, D = dn * inv_l;
    for (int64_t u = e0; u < e1; ++u) {
      float kr[C], vr[C], o1[C], o2[C];
      load(k + u * ldk + col, kr);
      load(v + u * ldv + col, vr);
      const float be = beta(u);
      const float a = __expf(dot(qr, kr) - m) * inv_l;
      const float ds = scale * a * (be * dot(gr, vr) - D), av = scale * a * be;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dqr[c] = fmaf(ds, kr[c], dqr[c]);
        o1[c] = ds * qr[c];
        o2[c] = av * gr[c];
      }
      store(dk + u * lddk + col, o1);
      store(dv + u * lddv + col, o2);
    }
    store(dq + i * lddq + col, dqr);
  }
}

// dst = split-fp16 of (scale * src)^T in one pass: the operands of dW = dY^T X straight from the row-major fp32 tensors (instead of
// a scaled copy, an fp32 transpose and a split pass each).  src [rows, cols] -> a_style: dst_hi [cols, 2 * rows_pad] with hi | lo in
// one row (A operand); else dst_hi, dst_lo [cols, rows_pad] (W operand).  Columns rows .. rows_pad are written as zeros.
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int64_t cols,
                                                              float scale, int64_t rows_pad, int a_style, __half* __restrict__ hi,
                                                              __half* __restrict__ lo, int vec) {
  // 64 x 64 tile: 256 B row segments in (float4 per lane when `vec`), 128 B column segments out (16 halves per lane)
  __shared__ float tile[64][65];
  const int64_t r0 = (int64_t)blockIdx.x * 64, c0 = (int64_t)blockIdx.y * 64;
  {
    const int c4 = (threadIdx.x & 15) * 4, rr = threadIdx.x >> 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rl = rr + 16 * j;
      const int64_t r = r0 + rl, c = c0 + c4;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (r < rows) {
        if (vec && c + 3 < cols) {
          const float4 v4 = *reinterpret_cast<const float4*>(src + r * ld_src + c);
          x[0] = v4.x; x[1] = v4.y; x[2] = v4.z; x[3] = v4.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c + e < cols) x[e] = src[r * ld_src + c + e];
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) tile[rl][c4 + e] = scale * x[e];
    }
  }
  __syncthreads();
  const int64_t ld = a_style ? 2 * rows_pad : rows_pad;
  __half* lo_base = a_style ? hi + rows_pad : lo;
  const int cl = threadIdx.x >> 2, rq = (threadIdx.x & 3) * 16;
  const int64_t c = c0 + cl, r = r0 + rq;
  if (c < cols && r < rows_pad) {
    __align__(16) __half2 h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x0 = fminf(fmaxf(tile[rq + 2 * e][cl], -65504.f), 65504.f), x1 = fminf(fmaxf(tile[rq + 2 * e + 1][cl], -65504.f), 65504.f);
      h[e] = __floats2half2_rn(x0, x1);
      const float2 f = __half22float2(h[e]);
      l[e] = __floats2half2_rn(x0 - f.x, x1 - f.y);
    }
    __half* ph = hi + c * ld + r;
    __half* pl = lo_base + c * ld + r;
    if (vec && r + 15 < rows_pad) {                                // rows_pad % 8 == 0 and 16 B aligned bases (checked by the host)
      reinterpret_cast<uint4*>(ph)[0] = reinterpret_cast<const uint4*>(h)[0];
      reinterpret_cast<uint4*>(ph)[1] = reinterpret_cast<const uint4*>(h)[1];
      reinterpret_cast<uint4*>(pl)[0] = reinterpret_cast<const uint4*>(l)[0];
      reinterpret_cast<uint4*>(pl)[1] = reinterpret_cast<const uint4*>(l)[1];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (r + 2 * e < rows_pad) {                                // rows_pad is even: pairs are written together
          reinterpret_cast<__half2*>(ph)[e] = h[e];
          reinterpret_cast<__half2*>(pl)[e] = l[e];
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------- training forward, fast forms
// (the dropout rates of transformer_lm_wiki103 make these the forward of every real training step; the generic one-warp-per-
// destination kernel above streams the 4.7 M causal edges of a Wiki103 block at 60 ms per layer)
//
// causal edges: the GEMM form of attn_gemm.cu with the dropout multiplier applied to the softmax weights -- S [H, L, L] from
// gnnlm_linear_batched_f16x3 (causal = 1) -> P~ = beta softmax_causal(S) as split fp16 [H, L, 2L] for the P~ V' product (causal = 2).
__global__ void __launch_bounds__(256) causal_softmax_drop_kernel(const float* __restrict__ S, int64_t L, int64_t ctx, int H, int64_t k_tile,
                                                                  int64_t row0, __half* __restrict__ P, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)H * L;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const int64_t i = r % L;
    const int head = (int)(r / L);
    const float* s = S + r * L;
    __half* p = P + r * 2 * L;
    const int64_t lo_j = (ctx > 0 && i + 1 > ctx) ? i + 1 - ctx : 0;
    float m = -INFINITY, l = 0.f;
    for (int64_t j = (lo_j & ~(int64_t)3) + lane * 4; j <= i; j += 128) {
      const float4 x4 = *reinterpret_cast<const float4*>(s + j);
      float x[4] = {x4.x, x4.y, x4.z, x4.w};
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (j + e < lo_j || j + e > i) x[e] = -INFINITY;
        mx = fmaxf(mx, x[e]);
      }
      if (mx > m) {
        l *= __expf(m - mx);
        m = mx;
      }
      if (m > -INFINITY) {
#pragma unroll
        for (int e = 0; e < 4; ++e) l += __expf(x[e] - m);
      }
    }
    const float M = warp_max(m);
    l = warp_sum(m > -INFINITY ? l * __expf(m - M) : 0.f);
    const float inv = 1.f / l;
    const int64_t j_end = k_tile > 0 ? min(L, (i / k_tile + 1) * k_tile) : L;
    for (int64_t j = lane * 4; j < j_end; j += 128) {
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      if (j <= i && j + 3 >= lo_j) {
        const float4 x4 = *reinterpret_cast<const float4*>(s + j);
        const float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j + e >= lo_j && j + e <= i)
            w[e] = __expf(x[e] - M) * inv *
                   (ad.p_thresh ? dm_scale(ad.seed, dm_edge(row0 + i, row0 + j + e, head), ad.p_thresh, ad.keep_scale) : 1.f);
      }
      uint2 hi, lo;
      split4_f16(w[0], w[1], w[2], w[3], hi, lo);
      *reinterpret_cast<uint2*>(p + j) = hi;
      *reinterpret_cast<uint2*>(p + L + j) = lo;
    }
  }
}

// ntgt-intra-ntgt chains (forward of cluster_attn_bwd_kernel): one warp per (cluster, head), three-row window.
template <int C>
__global__ void __launch_bounds__(256) cluster_attn_train_fwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                                     int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                                     const int32_t* __restrict__ node_base,
                                                                     const int32_t* __restrict__ cluster_nl, int64_t n_clusters, int H,
                                                                     float scale, float* __restrict__ out, int64_t ldo, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int64_t n_items = n_clusters * H;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  auto load = [&](const float* __restrict__ p, float (&r)[C]) {
    if constexpr (C == 4) {
      const float4 x = *reinterpret_cast<const float4*>(p);
      r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) r[c] = p[c];
    }
  };
  auto dot = [&](const float (&a)[C], const float (&b)[C]) {
    float p = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) p = fmaf(a[c], b[c], p);
    return warp_sum(p);
  };
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += warps) {
    const int64_t cl = it / H;
    const int head = (int)(it - cl * H);
    const int base = __ldg(node_base + cl);
    const int w = __ldg(node_base + cl + 1) - base;
    if (w <= 0) continue;
    const int nl = __ldg(cluster_nl + cl);
    const int col = head * 32 * C + lane * C;
    auto id_of = [&](int p) { return (int64_t)base + (p == nl ? 0 : (p < nl ? p + 1 : p)); };
    float Kp[C], Vp[C], Kc[C], Vc[C], Kn[C], Vn[C];
#pragma unroll
    for (int c = 0; c < C; ++c) Kp[c] = Vp[c] = Kn[c] = Vn[c] = 0.f;
    int64_t id_p = 0, id_c = id_of(0), id_n = 0;
    load(k + id_c * ldk + col, Kc);
    load(v + id_c * ldv + col, Vc);
    for (int p = 0; p < w; ++p) {
      const bool has_l = p > 0, has_r = p + 1 < w;
      if (has_r) {
        id_n = id_of(p + 1);
        load(k + id_n * ldk + col, Kn);
        load(v + id_n * ldv + col, Vn);
      }
      float qr[C];
      load(q + id_c * ldq + col, qr);
      const float s1 = dot(qr, Kc), s0 = has_l ? dot(qr, Kp) : -INFINITY, s2 = has_r ? dot(qr, Kn) : -INFINITY;
      const float m = fmaxf(s1, fmaxf(s0, s2));
      const float e0 = has_l ? __expf(s0 - m) : 0.f, e1 = __expf(s1 - m), e2 = has_r ? __expf(s2 - m) : 0.f;
      const float inv = scale / (e0 + e1 + e2);
      auto beta = [&](int64_t u) { return ad.p_thresh ? dm_scale(ad.seed, dm_edge(id_c, u, head), ad.p_thresh, ad.keep_scale) : 1.f; };
      const float a0 = has_l ? e0 * inv * beta(id_p) : 0.f, a1 = e1 * inv * beta(id_c), a2 = has_r ? e2 * inv * beta(id_n) : 0.f;
      float o[C];
#pragma unroll
      for (int c = 0; c < C; ++c) o[c] = a0 * Vp[c] + a1 * Vc[c] + a2 * Vn[c];
      float* op = out + id_c * ldo + col;
      if constexpr (C == 4) {
        *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) op[c] = o[c];
      }
      id_p = id_c; id_c = id_n;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        Kp[c] = Kc[c]; Vp[c] = Vc[c];
        Kc[c] = Kn[c]; Vc[c] = Vn[c];
      }
    }
  }
}

// contiguous sources per destination (indices == NULL: the inter edges): one warp per (destination, head), ONE pass with an
// online softmax -- out = scale * (sum_e w_e beta_e V'_e) / (sum_e w_e): dropout multiplies the normalised weights.
template <int C>
__global__ void __launch_bounds__(256) ranged_attn_train_fwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                                    int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                                    const int32_t* __restrict__ indptr, int64_t n_dst, int H, float scale,
                                                                    int accumulate, float* __restrict__ out, int64_t ldo, AttnDrop ad) {
  const int lane = threadIdx.x & 31;
  const int64_t n_items = n_dst * H;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  auto load = [&](const float* __restrict__ p, float (&r)[C]) {
    if constexpr (C == 4) {
      const float4 x = *reinterpret_cast<const float4*>(p);
      r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) r[c] = p[c];
    }
  };
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += warps) {
    const int64_t i = it / H;
    const int head = (int)(it - i * H);
    const int col = head * 32 * C + lane * C;
    const int64_t e0 = __ldg(indptr + i), e1 = __ldg(indptr + i + 1);
    float qr[C], acc[C];
    load(q + i * ldq + col, qr);
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int64_t u = e0; u < e1; ++u) {
      float kr[C], vr[C];
      load(k + u * ldk + col, kr);
      load(v + u * ldv + col, vr);
      float sc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) sc = fmaf(qr[c], kr[c], sc);
      sc = warp_sum(sc);
      const float mx = fmaxf(m, sc), corr = __expf(m - mx), wgt = __expf(sc - mx);
      const float wb = wgt * (ad.p_thresh ? dm_scale(ad.seed, dm_edge(i, u, head), ad.p_thresh, ad.keep_scale) : 1.f);
      l = l * corr + wgt;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(wb, vr[c], acc[c] * corr);
      m = mx;
    }
    const float f = l > 0.f ? scale / l : 0.f;
    float* op = out + i * ldo + col;
#pragma unroll
    for (int c = 0; c < C; ++c) op[c] = (accumulate ? op[c] : 0.f) + f * acc[c];
  }
}

}  // namespace gnnlm

using namespace gnnlm;

#define BWD_DISPATCH_C(Cv, ...)                                                                   \
  switch (Cv) {                                                                                   \
    case 1: edge_attn_bwd_kernel<1><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                 \
    case 2: edge_attn_bwd_kernel<2><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                 \
    case 4: edge_attn_bwd_kernel<4><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                 \
    case 8: edge_attn_bwd_kernel<8><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                 \
    case 16: edge_attn_bwd_kernel<16><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;               \
    default: edge_attn_bwd_kernel<32><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;               \
  }

extern "C" int32_t gnnlm_hgt_edge_attn_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                           const float* dout, int64_t ldo, const int32_t* indptr, const int32_t* indices,
                                           const int32_t* dst_ids, int64_t n_dst_cap, const int32_t* n_dst_dev, int64_t causal_L,
                                           int64_t intra_ctx, int32_t H, int32_t d_k, float scale, float* dq, int64_t lddq, float* dk,
                                           int64_t lddk, float* dv, int64_t lddv, float p_drop, uint64_t seed, cudaStream_t stream) {
  GNNLM_CHECK_ARG(q && k && v && dout && dq && dk && dv && (causal_L > 0 || indptr), GNNLM_E_ARG, "gnnlm_hgt_edge_attn_bwd: null pointer");
  GNNLM_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, GNNLM_E_ARG, "gnnlm_hgt_edge_attn_bwd: dropout rate must be in [0, 1)");
  const AttnDrop ad{seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop)};
  const int64_t d = (int64_t)H * d_k;
  GNNLM_CHECK_ARG(H > 0 && d_k > 0 && d % 32 == 0 && 32 % H == 0 && d <= 1024, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_edge_attn_bwd: needs d %% 32 == 0, d <= 1024 and H dividing 32 (H=%d d_k=%d)", H, d_k);
  const int C = (int)(d / 32);
  GNNLM_CHECK_ARG(C == 1 || C == 2 || C == 4 || C == 8 || C == 16 || C == 32, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_edge_attn_bwd: d / 32 must be a power of two");
  if (n_dst_cap == 0) return 0;
  const int group = 32 / H;
  const unsigned grid = (unsigned)(ceil_div(n_dst_cap, 8) < 148 * 32 ? ceil_div(n_dst_cap, 8) : 148 * 32);
  const int exclusive = causal_L == 0 && indices == nullptr;      // source id = edge id: no two edges share a source
  {
    const bool aligned = ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && lddq % 4 == 0 && lddk % 4 == 0 && lddv % 4 == 0 &&
                         ((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) % 16 == 0;
    static const bool off = [] { const char* e = getenv("GNNLM_TRAIN_RANGED_BWD"); return e && e[0] == '0'; }();   // A/B switch
    if (exclusive && !off && !dst_ids && !n_dst_dev && n_dst_cap > 0 && (d_k == 32 || d_k == 64 || (d_k == 128 && aligned))) {
      int64_t blocks = ceil_div(n_dst_cap * H, 8);
      if (blocks > 148 * 64) blocks = 148 * 64;
#define GNNLM_RAB(Cv) ranged_attn_bwd_kernel<Cv><<<(unsigned)blocks, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, dout, ldo, indptr, n_dst_cap, H, scale, \
                                                                                dq, lddq, dk, lddk, dv, lddv, ad)
      if (d_k == 128) GNNLM_RAB(4);
      else if (d_k == 64) GNNLM_RAB(2);
      else GNNLM_RAB(1);
#undef GNNLM_RAB
      GNNLM_LAUNCH_CHECK("gnnlm_hgt_edge_attn_bwd (ranged)");
      return 0;
    }
  }
  BWD_DISPATCH_C(C, q, ldq, k, ldk, v, ldv, dout, ldo, indptr, indices, dst_ids, n_dst_cap, n_dst_dev, causal_L, intra_ctx, group, scale,
                 dq, lddq, dk, lddk, dv, lddv, ad, (float*)nullptr, H, exclusive)
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_edge_attn_bwd");
  return 0;
}

extern "C" int32_t gnnlm_hgt_edge_attn_bwd_sym(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                               const float* dout, int64_t ldo, const int32_t* indptr, const int32_t* indices, int64_t n,
                                               int32_t H, int32_t d_k, float scale, float* dq, int64_t lddq, float* dk, int64_t lddk,
                                               float* dv, int64_t lddv, float* stats, float p_drop, uint64_t seed, cudaStream_t stream) {
  GNNLM_CHECK_ARG(q && k && v && dout && dq && dk && dv && stats && indptr && indices, GNNLM_E_ARG, "gnnlm_hgt_edge_attn_bwd_sym: null pointer");
  GNNLM_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, GNNLM_E_ARG, "gnnlm_hgt_edge_attn_bwd_sym: dropout rate must be in [0, 1)");
  const AttnDrop ad{seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop)};
  const int64_t d = (int64_t)H * d_k;
  GNNLM_CHECK_ARG(H > 0 && d_k > 0 && d % 32 == 0 && 32 % H == 0 && d <= 1024, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_edge_attn_bwd_sym: needs d %% 32 == 0, d <= 1024 and H dividing 32 (H=%d d_k=%d)", H, d_k);
  const int C = (int)(d / 32);
  GNNLM_CHECK_ARG(C == 1 || C == 2 || C == 4 || C == 8 || C == 16 || C == 32, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_edge_attn_bwd_sym: d / 32 must be a power of two");
  GNNLM_CHECK_ARG(n >= 0, GNNLM_E_SHAPE, "gnnlm_hgt_edge_attn_bwd_sym: bad sizes");
  if (n == 0) return 0;
  const int group = 32 / H;
  const unsigned grid = (unsigned)(ceil_div(n, 8) < 148 * 32 ? ceil_div(n, 8) : 148 * 32);
  const int32_t* no_ids = nullptr;
  BWD_DISPATCH_C(C, q, ldq, k, ldk, v, ldv, dout, ldo, indptr, indices, no_ids, n, no_ids, (int64_t)0, (int64_t)0, group, scale, dq, lddq,
                 dk, lddk, dv, lddv, ad, stats, H, 0)
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_edge_attn_bwd_sym (dq)");
  switch (C) {
    case 1: csr_bwd_dkv_kernel<1><<<grid, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, dout, ldo, indptr, indices, n, group, H, scale, stats, dk, lddk, dv, lddv, ad); break;
    case 2: csr_bwd_dkv_kernel<2><<<grid, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, dout, ldo, indptr, indices, n, group, H, scale, stats, dk, lddk, dv, lddv, ad); break;
    case 4: csr_bwd_dkv_kernel<4><<<grid, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, dout, ldo, indptr, indices, n, group, H, scale, stats, dk, lddk, dv, lddv, ad); break;
    case 8: csr_bwd_dkv_kernel<8><<<grid, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, dout, ldo, indptr, indices, n, group, H, scale, stats, dk, lddk, dv, lddv, ad); break;
    case 16: csr_bwd_dkv_kernel<16><<<grid, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, dout, ldo, indptr, indices, n, group, H, scale, stats, dk, lddk, dv, lddv, ad); break;
    default: csr_bwd_dkv_kernel<32><<<grid, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, dout, ldo, indptr, indices, n, group, H, scale, stats, dk, lddk, dv, lddv, ad); break;
  }
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_edge_attn_bwd_sym (dk, dv)");
  return 0;
}

#define CAUSAL_BWD_DISPATCH(KERNEL, Cv, ...)                                                       \
  switch (Cv) {                                                                                   \
    case 1: KERNEL<1><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                              \
    case 2: KERNEL<2><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                              \
    case 4: KERNEL<4><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                              \
    case 8: KERNEL<8><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                              \
    case 16: KERNEL<16><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                            \
    default: KERNEL<32><<<grid, 256, 0, stream>>>(__VA_ARGS__); break;                            \
  }

extern "C" int32_t gnnlm_hgt_causal_attn_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                             const float* dout, int64_t ldo, int64_t B, int64_t L, int64_t intra_ctx, int32_t H,
                                             int32_t d_k, float scale, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv,
                                             int64_t lddv, float* stats, float p_drop, uint64_t seed, cudaStream_t stream) {
  GNNLM_CHECK_ARG(q && k && v && dout && dq && dk && dv && stats, GNNLM_E_ARG, "gnnlm_hgt_causal_attn_bwd: null pointer");
  GNNLM_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, GNNLM_E_ARG, "gnnlm_hgt_causal_attn_bwd: dropout rate must be in [0, 1)");
  const AttnDrop ad{seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop)};
  const int64_t d = (int64_t)H * d_k;
  GNNLM_CHECK_ARG(H > 0 && d_k > 0 && d % 32 == 0 && 32 % H == 0 && d <= 1024, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_causal_attn_bwd: needs d %% 32 == 0, d <= 1024 and H dividing 32 (H=%d d_k=%d)", H, d_k);
  const int C = (int)(d / 32);
  GNNLM_CHECK_ARG(C == 1 || C == 2 || C == 4 || C == 8 || C == 16 || C == 32, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_causal_attn_bwd: d / 32 must be a power of two");
  GNNLM_CHECK_ARG(B >= 0 && L > 0, GNNLM_E_SHAPE, "gnnlm_hgt_causal_attn_bwd: bad sizes");
  const int64_t T = B * L;
  if (T == 0) return 0;
  const int group = 32 / H;
  const unsigned grid = (unsigned)(ceil_div(T, 8) < 148 * 32 ? ceil_div(T, 8) : 148 * 32);
  CAUSAL_BWD_DISPATCH(causal_bwd_dq_kernel, C, q, ldq, k, ldk, v, ldv, dout, ldo, T, L, intra_ctx, group, H, scale, dq, lddq, stats, ad)
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_causal_attn_bwd(dq)");
  CAUSAL_BWD_DISPATCH(causal_bwd_dkv_kernel, C, q, ldq, k, ldk, v, ldv, dout, ldo, T, L, intra_ctx, group, H, scale, stats, dk, lddk, dv,
                      lddv, ad)
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_causal_attn_bwd(dkv)");
  return 0;
}

extern "C" int32_t gnnlm_hgt_edge_attn_train_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                                 const int32_t* indptr, const int32_t* indices, int64_t n_dst, int64_t causal_L,
                                                 int64_t intra_ctx, int32_t H, int32_t d_k, float scale, int32_t accumulate, float* out,
                                                 int64_t ldo, float p_drop, uint64_t seed, cudaStream_t stream) {
  GNNLM_CHECK_ARG(q && k && v && out && (causal_L > 0 || indptr), GNNLM_E_ARG, "gnnlm_hgt_edge_attn_train_fwd: null pointer");
  GNNLM_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, GNNLM_E_ARG, "gnnlm_hgt_edge_attn_train_fwd: dropout rate must be in [0, 1)");
  const int64_t d = (int64_t)H * d_k;
  GNNLM_CHECK_ARG(H > 0 && d_k > 0 && d % 32 == 0 && 32 % H == 0 && d <= 1024, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_edge_attn_train_fwd: needs d %% 32 == 0, d <= 1024 and H dividing 32 (H=%d d_k=%d)", H, d_k);
  const int C = (int)(d / 32);
  GNNLM_CHECK_ARG(C == 1 || C == 2 || C == 4 || C == 8 || C == 16 || C == 32, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_edge_attn_train_fwd: d / 32 must be a power of two");
  if (n_dst == 0) return 0;
  const AttnDrop ad{seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop)};
  if (causal_L == 0 && indices == nullptr && (d_k == 32 || d_k == 64 || d_k == 128)) {         // contiguous sources: one pass per (dst, head)
    const bool aligned = ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) % 16 == 0;
    if (d_k != 128 || aligned) {
      int64_t blocks = ceil_div(n_dst * H, 8);
      if (blocks > 148 * 64) blocks = 148 * 64;
#define GNNLM_RAF(Cv) ranged_attn_train_fwd_kernel<Cv><<<(unsigned)blocks, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, indptr, n_dst, H, scale, \
                                                                                      accumulate, out, ldo, ad)
      if (d_k == 128) GNNLM_RAF(4);
      else if (d_k == 64) GNNLM_RAF(2);
      else GNNLM_RAF(1);
#undef GNNLM_RAF
      GNNLM_LAUNCH_CHECK("gnnlm_hgt_edge_attn_train_fwd (ranged)");
      return 0;
    }
  }
  const int group = 32 / H;
  const unsigned grid = (unsigned)(ceil_div(n_dst, 8) < 148 * 32 ? ceil_div(n_dst, 8) : 148 * 32);
  CAUSAL_BWD_DISPATCH(edge_attn_train_fwd_kernel, C, q, ldq, k, ldk, v, ldv, indptr, indices, n_dst, causal_L, intra_ctx, group, scale,
                      accumulate, out, ldo, ad)
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_edge_attn_train_fwd");
  return 0;
}

extern "C" int32_t gnnlm_dropout_f32(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t rows, int64_t cols, float p_drop,
                                     uint64_t seed, cudaStream_t stream) {
  GNNLM_CHECK_ARG(x && y, GNNLM_E_ARG, "gnnlm_dropout_f32: null pointer");
  GNNLM_CHECK_ARG(rows >= 0 && cols > 0 && ldx >= cols && ldy >= cols && p_drop >= 0.f && p_drop < 1.f, GNNLM_E_SHAPE,
                  "gnnlm_dropout_f32: bad shape / rate");
  if (rows == 0) return 0;
  const int64_t n = rows * cols;
  const unsigned grid = (unsigned)(ceil_div(n, 256) < 148 * 16 ? ceil_div(n, 256) : 148 * 16);
  dropout_kernel<<<grid, 256, 0, stream>>>(x, ldx, y, ldy, rows, cols, seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop));
  GNNLM_LAUNCH_CHECK("gnnlm_dropout_f32");
  return 0;
}

extern "C" int32_t gnnlm_layernorm_bwd(const float* o, int64_t ldo, const float* residual, int64_t ldr, const float* gamma, float eps,
                                       const float* dy, int64_t ldy, int64_t rows, const int32_t* rows_dev, int64_t d, float* dx,
                                       int64_t ldx, float* dgamma, float* dbeta, cudaStream_t stream) {
  GNNLM_CHECK_ARG(o && gamma && dy && dx && dgamma && dbeta, GNNLM_E_ARG, "gnnlm_layernorm_bwd: null pointer");
  GNNLM_CHECK_ARG(d > 0 && d <= 2048 && rows >= 0, GNNLM_E_SHAPE, "gnnlm_layernorm_bwd: d must be in (0, 2048]");
  if (rows == 0) return 0;
  const unsigned grid = (unsigned)(rows < 148 * 8 ? rows : 148 * 8);
  layernorm_bwd_kernel<<<grid, 256, 0, stream>>>(o, ldo, residual, ldr, gamma, eps, dy, ldy, rows, rows_dev, (int)d, dx, ldx, dgamma, dbeta);
  GNNLM_LAUNCH_CHECK("gnnlm_layernorm_bwd");
  return 0;
}

extern "C" int32_t gnnlm_xent_fwd_bwd(float* logits, int64_t ld, const int64_t* target, int64_t rows, int64_t C, float grad_scale,
                                      double* loss, cudaStream_t stream) {
  GNNLM_CHECK_ARG(logits && target && loss, GNNLM_E_ARG, "gnnlm_xent_fwd_bwd: null pointer");
  GNNLM_CHECK_ARG(rows >= 0 && C > 0 && ld >= C, GNNLM_E_SHAPE, "gnnlm_xent_fwd_bwd: bad shape");
  if (rows == 0) return 0;
  const unsigned grid = (unsigned)(rows < 148 * 8 ? rows : 148 * 8);
  xent_fwd_bwd_kernel<<<grid, 256, 0, stream>>>(logits, ld, target, rows, C, grad_scale, loss);
  GNNLM_LAUNCH_CHECK("gnnlm_xent_fwd_bwd");
  return 0;
}

extern "C" int32_t gnnlm_transpose_f32(const float* src, int64_t ld_src, int64_t rows, const int32_t* rows_dev, int64_t cols, float* dst,
                                       int64_t ld_dst, int64_t rows_pad, cudaStream_t stream) {
  GNNLM_CHECK_ARG(src && dst, GNNLM_E_ARG, "gnnlm_transpose_f32: null pointer");
  GNNLM_CHECK_ARG(rows >= 0 && cols > 0 && ld_src >= cols && rows_pad >= rows && ld_dst >= rows_pad, GNNLM_E_SHAPE,
                  "gnnlm_transpose_f32: bad shape");
  if (rows_pad == 0) return 0;
  dim3 grid((unsigned)ceil_div(rows_pad, 32), (unsigned)ceil_div(cols, 32));
  GNNLM_CHECK_ARG(grid.y < 65536, GNNLM_E_SHAPE, "gnnlm_transpose_f32: too many columns");
  transpose_kernel<<<grid, 256, 0, stream>>>(src, ld_src, rows, rows_dev, cols, dst, ld_dst, rows_pad);
  GNNLM_LAUNCH_CHECK("gnnlm_transpose_f32");
  return 0;
}

extern "C" int32_t gnnlm_colsum_f32(const float* x, int64_t ld, int64_t rows, const int32_t* rows_dev, int64_t cols, float* out,
                                    cudaStream_t stream) {
  GNNLM_CHECK_ARG(x && out, GNNLM_E_ARG, "gnnlm_colsum_f32: null pointer");
  GNNLM_CHECK_ARG(rows >= 0 && cols > 0 && ld >= cols, GNNLM_E_SHAPE, "gnnlm_colsum_f32: bad shape");
  if (rows == 0) return 0;
  dim3 grid((unsigned)ceil_div(cols, 256), (unsigned)(rows < 512 ? rows : 512));
  colsum_kernel<<<grid, 256, 0, stream>>>(x, ld, rows, rows_dev, cols, out);
  GNNLM_LAUNCH_CHECK("gnnlm_colsum_f32");
  return 0;
}

extern "C" int32_t gnnlm_axpy_f32(float* y, int64_t ldy, const float* x, int64_t ldx, int64_t rows, const int32_t* rows_dev, int64_t cols,
                                  float a, cudaStream_t stream) {
  GNNLM_CHECK_ARG(y && x, GNNLM_E_ARG, "gnnlm_axpy_f32: null pointer");
  GNNLM_CHECK_ARG(rows >= 0 && cols > 0 && ldy >= cols && ldx >= cols, GNNLM_E_SHAPE, "gnnlm_axpy_f32: bad shape");
  if (rows == 0) return 0;
  const int64_t n = rows * cols;
  const unsigned grid = (unsigned)(ceil_div(n, 256) < 148 * 16 ? ceil_div(n, 256) : 148 * 16);
  axpy_kernel<<<grid, 256, 0, stream>>>(y, ldy, x, ldx, rows, rows_dev, cols, a);
  GNNLM_LAUNCH_CHECK("gnnlm_axpy_f32");
  return 0;
}

extern "C" int32_t gnnlm_scatter_add_rows(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, const int32_t* ids, int64_t rows,
                                          const int32_t* rows_dev, int64_t cols, cudaStream_t stream) {
  GNNLM_CHECK_ARG(dst && src && ids, GNNLM_E_ARG, "gnnlm_scatter_add_rows: null pointer");
  GNNLM_CHECK_ARG(rows >= 0 && cols > 0 && ld_dst >= cols && ld_src >= cols, GNNLM_E_SHAPE, "gnnlm_scatter_add_rows: bad shape");
  if (rows == 0) return 0;
  const int64_t n = rows * cols;
  const unsigned grid = (unsigned)(ceil_div(n, 256) < 148 * 16 ? ceil_div(n, 256) : 148 * 16);
  scatter_add_rows_kernel<<<grid, 256, 0, stream>>>(dst, ld_dst, src, ld_src, ids, rows, rows_dev, cols);
  GNNLM_LAUNCH_CHECK("gnnlm_scatter_add_rows");
  return 0;
}

extern "C" int32_t gnnlm_causal_softmax_bwd_split(const float* S, const float* G, int64_t L, int64_t intra_ctx, int32_t H, int64_t row0,
                                                  float scale, float p_drop, uint64_t seed, float* stats, void* dS, void* WT, void* dST,
                                                  cudaStream_t stream) {
  GNNLM_CHECK_ARG(S && G && stats && dS && WT && dST, GNNLM_E_ARG, "gnnlm_causal_softmax_bwd_split: null pointer");
  GNNLM_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, GNNLM_E_ARG, "gnnlm_causal_softmax_bwd_split: dropout rate must be in [0, 1)");
  GNNLM_CHECK_ARG(L > 0 && L % CBT == 0 && L / CBT <= 65535 && H > 0 && H <= 65535 && row0 >= 0, GNNLM_E_SHAPE,
                  "gnnlm_causal_softmax_bwd_split: L must be a multiple of %d", CBT);
  GNNLM_CHECK_ARG((uintptr_t)S % 16 == 0 && (uintptr_t)G % 16 == 0 && (uintptr_t)dS % 8 == 0 && (uintptr_t)WT % 8 == 0 && (uintptr_t)dST % 8 == 0,
                  GNNLM_E_SHAPE, "gnnlm_causal_softmax_bwd_split: S / G must be 16 B, the split outputs 8 B aligned");
  const AttnDrop ad{seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop)};
  int64_t blocks = ceil_div((int64_t)H * L, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  causal_bwd_stats_kernel<<<(unsigned)blocks, 256, 0, stream>>>(S, G, L, intra_ctx, H, row0, stats, ad);
  GNNLM_LAUNCH_CHECK("gnnlm_causal_softmax_bwd_split (stats)");
  const dim3 grid((unsigned)(L / CBT), (unsigned)(L / CBT), (unsigned)H);
  causal_bwd_tile_kernel<<<grid, 256, 0, stream>>>(S, G, stats, L, intra_ctx, row0, scale, (__half*)dS, (__half*)WT, (__half*)dST, ad);
  GNNLM_LAUNCH_CHECK("gnnlm_causal_softmax_bwd_split (tiles)");
  return 0;
}

extern "C" int32_t gnnlm_hgt_cluster_attn_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                              const float* dout, int64_t ldo, const int32_t* node_base, const int32_t* cluster_nl,
                                              int64_t n_clusters, int32_t H, int32_t d_k, float scale, float* dq, int64_t lddq, float* dk,
                                              int64_t lddk, float* dv, int64_t lddv, float p_drop, uint64_t seed, cudaStream_t stream) {
  GNNLM_CHECK_ARG(q && k && v && dout && dq && dk && dv && node_base && cluster_nl, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_bwd: null pointer");
  GNNLM_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_bwd: dropout rate must be in [0, 1)");
  GNNLM_CHECK_ARG(H > 0 && (d_k == 32 || d_k == 64 || d_k == 128), GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_cluster_attn_bwd: d_k must be 32, 64 or 128 (one warp per cluster and head)");
  GNNLM_CHECK_ARG(d_k != 128 || (ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && lddq % 4 == 0 && lddk % 4 == 0 && lddv % 4 == 0 &&
                                 ((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) % 16 == 0),
                  GNNLM_E_SHAPE, "gnnlm_hgt_cluster_attn_bwd: rows must be 16 B aligned");
  GNNLM_CHECK_ARG(n_clusters >= 0, GNNLM_E_SHAPE, "gnnlm_hgt_cluster_attn_bwd: bad sizes");
  if (n_clusters == 0) return 0;
  const AttnDrop ad{seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop)};
  int64_t blocks = ceil_div(n_clusters * H, 8);
  if (blocks > 148 * 64) blocks = 148 * 64;
#define GNNLM_CAB(Cv) cluster_attn_bwd_kernel<Cv><<<(unsigned)blocks, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, dout, ldo, node_base, cluster_nl, \
                                                                                 n_clusters, H, scale, dq, lddq, dk, lddk, dv, lddv, ad)
  if (d_k == 128) GNNLM_CAB(4);
  else if (d_k == 64) GNNLM_CAB(2);
  else GNNLM_CAB(1);
#undef GNNLM_CAB
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_cluster_attn_bwd");
  return 0;
}

extern "C" int32_t gnnlm_transpose_split_f16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float scale, int64_t rows_pad,
                                             int32_t a_style, void* hi, void* lo, cudaStream_t stream) {
  GNNLM_CHECK_ARG(src && hi && (a_style || lo), GNNLM_E_ARG, "gnnlm_transpose_split_f16: null pointer");
  GNNLM_CHECK_ARG(rows >= 0 && cols > 0 && ld_src >= cols && rows_pad >= rows && rows_pad % 2 == 0 && (uintptr_t)hi % 4 == 0 &&
                      (a_style || (uintptr_t)lo % 4 == 0),
                  GNNLM_E_SHAPE, "gnnlm_transpose_split_f16: rows_pad must be even and >= rows");
  if (rows_pad == 0) return 0;
  const dim3 grid((unsigned)ceil_div(rows_pad, 64), (unsigned)ceil_div(cols, 64));
  const int vec = ld_src % 4 == 0 && (uintptr_t)src % 16 == 0 && rows_pad % 8 == 0 && (uintptr_t)hi % 16 == 0 && (a_style || (uintptr_t)lo % 16 == 0);
  transpose_split_kernel<<<grid, 256, 0, stream>>>(src, ld_src, rows, cols, scale, rows_pad, a_style, (__half*)hi, (__half*)lo, vec);
  GNNLM_LAUNCH_CHECK("gnnlm_transpose_split_f16");
  return 0;
}

extern "C" int32_t gnnlm_causal_softmax_drop_split(const float* S, int64_t L, int64_t intra_ctx, int32_t H, int64_t k_tile, int64_t row0,
                                                   float p_drop, uint64_t seed, void* P, cudaStream_t stream) {
  GNNLM_CHECK_ARG(S && P, GNNLM_E_ARG, "gnnlm_causal_softmax_drop_split: null pointer");
  GNNLM_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, GNNLM_E_ARG, "gnnlm_causal_softmax_drop_split: dropout rate must be in [0, 1)");
  GNNLM_CHECK_ARG(L > 0 && L % 4 == 0 && H > 0 && k_tile >= 0 && row0 >= 0 && (uintptr_t)S % 16 == 0 && (uintptr_t)P % 8 == 0, GNNLM_E_SHAPE,
                  "gnnlm_causal_softmax_drop_split: L must be a multiple of 4 and S / P 16 B / 8 B aligned");
  const AttnDrop ad{seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop)};
  int64_t blocks = ceil_div((int64_t)H * L, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  causal_softmax_drop_kernel<<<(unsigned)blocks, 256, 0, stream>>>(S, L, intra_ctx, H, k_tile, row0, (__half*)P, ad);
  GNNLM_LAUNCH_CHECK("gnnlm_causal_softmax_drop_split");
  return 0;
}

extern "C" int32_t gnnlm_hgt_cluster_attn_train_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                                    const int32_t* node_base, const int32_t* cluster_nl, int64_t n_clusters, int32_t H,
                                                    int32_t d_k, float scale, float* out, int64_t ldo, float p_drop, uint64_t seed,
                                                    cudaStream_t stream) {
  GNNLM_CHECK_ARG(q && k && v && out && node_base && cluster_nl, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_train_fwd: null pointer");
  GNNLM_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_train_fwd: dropout rate must be in [0, 1)");
  GNNLM_CHECK_ARG(H > 0 && (d_k == 32 || d_k == 64 || d_k == 128), GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_cluster_attn_train_fwd: d_k must be 32, 64 or 128 (one warp per cluster and head)");
  GNNLM_CHECK_ARG(d_k != 128 || (ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 &&
                                 ((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0),
                  GNNLM_E_SHAPE, "gnnlm_hgt_cluster_attn_train_fwd: rows must be 16 B aligned");
  if (n_clusters <= 0) return 0;
  const AttnDrop ad{seed, (uint32_t)(p_drop * 16777216.f), 1.f / (1.f - p_drop)};
  int64_t blocks = ceil_div(n_clusters * H, 8);
  if (blocks > 148 * 64) blocks = 148 * 64;
#define GNNLM_CAF(Cv) cluster_attn_train_fwd_kernel<Cv><<<(unsigned)blocks, 256, 0, stream>>>(q, ldq, k, ldk, v, ldv, node_base, cluster_nl, \
                                                                                       n_clusters, H, scale, out, ldo, ad)
  if (d_k == 128) GNNLM_CAF(4);
  else if (d_k == 64) GNNLM_CAF(2);
  else GNNLM_CAF(1);
#undef GNNLM_CAF
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_cluster_attn_train_fwd");
  return 0;
}
