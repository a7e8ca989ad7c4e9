// ntgt-intra-ntgt attention without the CSR indirection (gnnlm_hgt_cluster_attn).
//
// Same arithmetic as gnnlm_hgt_edge_attn over (nn_indptr, nn_indices) -- i.e. the DGL sequence of
// HGTLayer.forward (reference: fairseq/models/hgt.py:350-358,383-386) for the ('ntgt','intra','ntgt')
// edge type -- but driven by the structure graph assembly already knows: every valid
// (token, neighbour) pair owns a cluster of w contiguous node ids whose intra edges are a chain with
// self loops (build_ntgt_edges(context=1, bidirect=True), fairseq/data/token_block_dataset.py:395-400).
//
// Why a second kernel: the CSR form is latency-bound (indptr -> indices -> rows is three dependent
// round trips per destination, and each K'/V' row is fetched by three destinations).  Here one warp owns
// one (cluster, feature slice): one round trip for (node_base, cluster_nl), then the 3w row segments of
// Q / K' / V' are all in flight together and each is read exactly once -- HBM traffic equals the
// algorithmic bytes of SURVEY.md 8(d): N*(2+1)*d*s read + N*d*4 written.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace gnnlm {

constexpr int CA_THREADS = 256;

__device__ __forceinline__ int ca_pos_to_id(int q, int nl) { return q == nl ? 0 : (q < nl ? q + 1 : q); }

// dot product of the lane's C features, reduced over the GROUP lanes that share a head (GROUP == 0: run-time `group`)
template <int C, int GROUP>
__device__ __forceinline__ float group_dot(const float (&a)[C], const float (&b)[C], int group) {
  float p = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) p = fmaf(a[c], b[c], p);
  if constexpr (GROUP > 0) {
#pragma unroll
    for (int o = GROUP >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  } else {
    for (int o = group >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  }
  return p;
}

// softmax over the (up to) three chain scores and the weighted sum of the matching V' rows
template <int C>
__device__ __forceinline__ void chain_mix(float s0, float s1, float s2, bool has_l, bool has_r, const float (&vp)[C],
                                          const float (&vc)[C], const float (&vn)[C], float (&r)[C]) {
  if (!has_l) s0 = -INFINITY;
  if (!has_r) s2 = -INFINITY;
  const float mx = fmaxf(s1, fmaxf(s0, s2));
  const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx), e2 = __expf(s2 - mx);      // exp(-inf) = 0 for a missing side
  const float inv = __fdividef(1.f, e0 + e1 + e2);
  const float w0 = e0 * inv, w1 = e1 * inv, w2 = e2 * inv;
#pragma unroll
  for (int cc = 0; cc < C; ++cc) {
    float a = w1 * vc[cc];
    if (has_l) a = fmaf(w0, vp[cc], a);
    if (has_r) a = fmaf(w2, vn[cc], a);
    r[cc] = a;
  }
}

// q8p (nullable): where the e4m3 companion (hi8 at q8p, lo8 at q8p + lo_off) of a split-fp16 output element group goes;
// write_lo = false: the fp16 lo half is not stored (hi + companion only: the operand set of gnnlm_linear_f16f8)
template <typename OutT, int C>
__device__ __forceinline__ void store_out(OutT* __restrict__ p, const float (&r)[C], int lo_off, uint8_t* q8p = nullptr,
                                          bool write_lo = true) {
  if constexpr (sizeof(OutT) == 4) {
    store_f32<C>(reinterpret_cast<float*>(p), r);
  } else if constexpr (std::is_same<OutT, __half>::value) {     // split-fp16: hi at p, lo at p + lo_off
    static_assert(C % 4 == 0, "split output needs 8 B chunks");
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      uint2 hi, lo;
      split4_f16(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3], hi, lo);
      reinterpret_cast<uint2*>(p)[i] = hi;
      if (write_lo) reinterpret_cast<uint2*>(p + lo_off)[i] = lo;
      if (q8p) {
        uint32_t h8, l8;
        q8_from_split4(hi, lo, h8, l8);
        reinterpret_cast<uint32_t*>(q8p)[i] = h8;
        reinterpret_cast<uint32_t*>(q8p + lo_off)[i] = l8;
      }
    }
  } else {
    static_assert(C % 8 == 0 || C == 4, "bf16 output needs 8 B / 16 B chunks");
    if constexpr (C == 4) {
      __nv_bfloat162 a = __floats2bfloat162_rn(r[0], r[1]), b = __floats2bfloat162_rn(r[2], r[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(p) = u;
    } else {
#pragma unroll
      for (int i = 0; i < C / 8; ++i) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(r[8 * i + 2 * j], r[8 * i + 2 * j + 1]);
        reinterpret_cast<uint4*>(p)[i] = *reinterpret_cast<uint4*>(h);
      }
    }
  }
}

// One warp per (cluster, 32*C-feature slice).  WMAX > 0: all nodes of clusters of at most WMAX (1, 3, 5, 7) nodes, every row
// loaded up front.  WMAX == 0: centre-only variant (its own instantiation: 3 rows of K'/V' per cluster and a third of the
// registers of the all-nodes form).  GROUP: lanes per head (0 = run time).
// HQ: q / k / v are GNNLM_F24 (T = a 2-byte element: the 16-bit plane; the byte of the element at 16-bit address a is at a / 2 + bias)
// (Measured on the Wiki103 shape: the centre-only form gains from the 3-byte rows, 0.47 -> 0.43 ms; the all-nodes form does not --
// 0.78 -> 0.93 ms even at four CTAs per SM: two narrower loads per row segment plus two byte-permute instructions per element make
// it issue-bound -- so the host only feeds GNNLM_F24 to the centre-only layer.)
template <typename T, typename OutT, int C, int WMAX, int GROUP, bool HQ = false>
__global__ void __launch_bounds__(CA_THREADS) cluster_attn_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k,
                                                                  int64_t ldk, const T* __restrict__ v, int64_t ldv,
                                                                  const int32_t* __restrict__ node_base,
                                                                  const int32_t* __restrict__ valid_base,
                                                                  const int32_t* __restrict__ cluster_nl, int64_t n_clusters,
                                                                  int group, int n_slices,
                                                                  OutT* __restrict__ out, int64_t ldo, int lo_off,
                                                                  uint8_t* __restrict__ q8, int64_t ldq8, bool write_lo,
                                                                  int64_t q_bias = 0, int64_t k_bias = 0, int64_t v_bias = 0,
                                                                  const float2* __restrict__ kv_stats = nullptr,
                                                                  const float* __restrict__ k_c = nullptr,
                                                                  const float* __restrict__ k_b = nullptr,
                                                                  const float* __restrict__ v_c = nullptr,
                                                                  const float* __restrict__ v_b = nullptr) {
  constexpr bool centre_only = WMAX == 0;
  auto ld_row = [](const T* __restrict__ p, int64_t bias, float (&r)[C]) {
    if constexpr (HQ) load_row_f24(p, bias, r);
    else load_row<T, C>(p, r);
  };
  constexpr int WM = WMAX > 0 ? WMAX : 1;
  const int lane = threadIdx.x & 31;
  const int64_t n_items = n_clusters * n_slices;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += warps) {
    const int64_t c = it / n_slices;
    const int col = ((int)(it - c * n_slices) * 32 + lane) * C;
    const int base = __ldg(node_base + c);
    const int w = __ldg(node_base + c + 1) - base;
    if (w <= 0) continue;                                    // invalid neighbour: no cluster
    const int nl = __ldg(cluster_nl + c);
    if constexpr (!centre_only) {
      // rows in sorted (chain) order: position p holds node id base + ca_pos_to_id(p, nl)
      const T* qb = q + (int64_t)base * ldq + col;
      const T* kb = k + (int64_t)base * ldk + col;
      const T* vb = v + (int64_t)base * ldv + col;
      float qq[WM][C], kk[WM][C], vv[WM][C];
#pragma unroll
      for (int p = 0; p < WM; ++p) {
        if (p < w) {
          const int64_t id = ca_pos_to_id(p, nl);
          ld_row(qb + id * ldq, q_bias, qq[p]);
          ld_row(kb + id * ldk, k_bias, kk[p]);
          ld_row(vb + id * ldv, v_bias, vv[p]);
        }
      }
#pragma unroll
      for (int p = 0; p < WM; ++p) {
        if (p < w) {
          constexpr int pl = 0;
          const int pp = p > 0 ? p - 1 : pl, pn = p + 1 < WM ? p + 1 : p;
          const bool has_l = p > 0, has_r = p + 1 < WM && p + 1 < w;
          const float s1 = group_dot<C, GROUP>(qq[p], kk[p], group);
          const float s0 = has_l ? group_dot<C, GROUP>(qq[p], kk[pp], group) : 0.f;
          const float s2 = p + 1 < WM ? group_dot<C, GROUP>(qq[p], kk[pn], group) : 0.f;
          float r[C];
          chain_mix<C>(s0, s1, s2, has_l, has_r, vv[pp], vv[p], vv[pn], r);
          const int64_t row = base + ca_pos_to_id(p, nl);
          store_out<OutT, C>(out + row * ldo + col, r, lo_off, q8 ? q8 + row * ldq8 + col : nullptr, write_lo);
        }
      }
    } else {
      // centre node (sorted position nl) attends to positions nl-1, nl, nl+1
      const int64_t ci = __ldg(valid_base + c);             // compact row of q / out
      float qq[C], kk[3][C], vv[3][C];
      ld_row(q + ci * ldq + col, q_bias, qq);
      const bool has_l = nl > 0, has_r = nl + 1 < w;
      const int64_t idl = base + (has_l ? ca_pos_to_id(nl - 1, nl) : 0), idr = base + (has_r ? ca_pos_to_id(nl + 1, nl) : 0);
      ld_row(k + (int64_t)base * ldk + col, k_bias, kk[1]);
      ld_row(v + (int64_t)base * ldv + col, v_bias, vv[1]);
      ld_row(k + idl * ldk + col, k_bias, kk[0]);            // an absent side re-reads the centre row (L1 hit), masked below
      ld_row(v + idl * ldv + col, v_bias, vv[0]);
      ld_row(k + idr * ldk + col, k_bias, kk[2]);
      ld_row(v + idr * ldv + col, v_bias, vv[2]);
      float r[C];
      bool done = false;
      if constexpr (HQ) {
        // deferred LayerNorm of the previous layer (rowops.cu rowstats_q8_kernel): the rows are RAW products z' W~^T and the K' / V'
        // of node m are rs_m (raw_m - mean_m c) + b with per-column vectors (c, b) and per-node statistics (mean_m, rs_m).  Applied
        // without touching the rows: <q, K'_m> = rs_m <q, k_raw_m> + <q, k_b> - mean_m rs_m <q, k_c>, and
        // sum_m w_m V'_m = sum_m (w_m rs_m) v_raw_m + v_b - v_c sum_m w_m mean_m rs_m   (the weights sum to one).
        if (kv_stats) {
          float2 st[3];
          st[0] = __ldg(kv_stats + idl);
          st[1] = __ldg(kv_stats + base);
          st[2] = __ldg(kv_stats + idr);
          float tmp[C];
          load_f32<C>(k_b + col, tmp);
          const float qb = group_dot<C, GROUP>(qq, tmp, group);
          load_f32<C>(k_c + col, tmp);
          const float qc = group_dot<C, GROUP>(qq, tmp, group);
          float sc[3], wm[3];
#pragma unroll
          for (int p = 0; p < 3; ++p) sc[p] = fmaf(st[p].y, group_dot<C, GROUP>(qq, kk[p], group), qb - st[p].x * st[p].y * qc);
          if (!has_l) sc[0] = -INFINITY;
          if (!has_r) sc[2] = -INFINITY;
          const float mx = fmaxf(sc[1], fmaxf(sc[0], sc[2]));
          const float e0 = __expf(sc[0] - mx), e1 = __expf(sc[1] - mx), e2 = __expf(sc[2] - mx);
          const float inv = __fdividef(1.f, e0 + e1 + e2);
          wm[0] = e0 * inv * st[0].y; wm[1] = e1 * inv * st[1].y; wm[2] = e2 * inv * st[2].y;      // w_m rs_m (0 for a missing side)
          const float shift = wm[0] * st[0].x + wm[1] * st[1].x + wm[2] * st[2].x;               // sum_m w_m rs_m mean_m
          float vb[C];
          load_f32<C>(v_c + col, tmp);
          load_f32<C>(v_b + col, vb);
#pragma unroll
          for (int cc = 0; cc < C; ++cc) {
            float a = fmaf(wm[1], vv[1][cc], fmaf(-shift, tmp[cc], vb[cc]));
            if (has_l) a = fmaf(wm[0], vv[0][cc], a);
            if (has_r) a = fmaf(wm[2], vv[2][cc], a);
            r[cc] = a;
          }
          done = true;
        }
      }
      if (!done) {
        const float s1 = group_dot<C, GROUP>(qq, kk[1], group);
        const float s0 = group_dot<C, GROUP>(qq, kk[0], group);
        const float s2 = group_dot<C, GROUP>(qq, kk[2], group);
        chain_mix<C>(s0, s1, s2, has_l, has_r, vv[0], vv[1], vv[2], r);
      }
      store_out<OutT, C>(out + ci * ldo + col, r, lo_off, q8 ? q8 + ci * ldq8 + col : nullptr, write_lo);
    }
  }
}

// Clusters of any size (used for w > 3): walk the chain in sorted order with a ring of four (Q, K', V') row slots --
// position p reads slots p-1, p, p+1 while the loads of position p+2 are in flight -- so the register count (and the
// number of resident warps) does not grow with w the way the fully unrolled form above does.  The ring is unrolled by
// four so that every slot index is a compile-time constant (no register moves).
template <typename T, typename OutT, int C, int GROUP>
__global__ void __launch_bounds__(CA_THREADS) cluster_attn_window_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k,
                                                                         int64_t ldk, const T* __restrict__ v, int64_t ldv,
                                                                         const int32_t* __restrict__ node_base,
                                                                         const int32_t* __restrict__ cluster_nl, int64_t n_clusters,
                                                                         int group, int n_slices, OutT* __restrict__ out,
                                                                         int64_t ldo, int lo_off, uint8_t* __restrict__ q8,
                                                                         int64_t ldq8, bool write_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t n_items = n_clusters * n_slices;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += warps) {
    const int64_t c = it / n_slices;
    const int col = ((int)(it - c * n_slices) * 32 + lane) * C;
    const int base = __ldg(node_base + c);
    const int w = __ldg(node_base + c + 1) - base;
    if (w <= 0) continue;
    const int nl = __ldg(cluster_nl + c);
    const T* qb = q + (int64_t)base * ldq + col;
    const T* kb = k + (int64_t)base * ldk + col;
    const T* vb = v + (int64_t)base * ldv + col;
    OutT* ob = out + (int64_t)base * ldo + col;
    float Q[4][C], K[4][C], V[4][C];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (p < w) {
        const int64_t id = ca_pos_to_id(p, nl);
        load_row<T, C>(qb + id * ldq, Q[p]);
        load_row<T, C>(kb + id * ldk, K[p]);
        load_row<T, C>(vb + id * ldv, V[p]);
      }
    }
    for (int p0 = 0; p0 < w; p0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u;
        if (p < w) {
          const int sp = (u + 3) & 3, sn = (u + 1) & 3, s2_ = (u + 2) & 3;       // constants once unrolled
          if (p + 2 < w) {                                      // in flight while position p is being computed
            const int64_t id = ca_pos_to_id(p + 2, nl);
            load_row<T, C>(qb + id * ldq, Q[s2_]);
            load_row<T, C>(kb + id * ldk, K[s2_]);
            load_row<T, C>(vb + id * ldv, V[s2_]);
          }
          const bool has_l = p > 0, has_r = p + 1 < w;
          const float s1 = group_dot<C, GROUP>(Q[u], K[u], group);
          const float s0 = group_dot<C, GROUP>(Q[u], K[sp], group);
          const float s2 = group_dot<C, GROUP>(Q[u], K[sn], group);
          float r[C];
          chain_mix<C>(s0, s1, s2, has_l, has_r, V[sp], V[u], V[sn], r);
          const int64_t rid = ca_pos_to_id(p, nl);
          store_out<OutT, C>(ob + rid * ldo, r, lo_off, q8 ? q8 + (base + rid) * ldq8 + col : nullptr, write_lo);
        }
      }
    }
  }
}

// GNNLM_CLUSTER_WINDOW=1: use the ring-of-four window kernel for every chain longer than 3 (A/B timing switch)
static bool window_forced() {
  static const bool f = [] { const char* e = getenv("GNNLM_CLUSTER_WINDOW"); return e && e[0] == '1'; }();
  return f;
}

template <typename T, typename OutT, int C, int GROUP>
static int32_t dispatch_w(int wmax, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          const int32_t* node_base, const int32_t* valid_base, const int32_t* cluster_nl, int64_t n_clusters,
                          int centre_only, int group, int n_slices, void* out, int64_t ldo, int lo_off, uint8_t* q8,
                          int64_t ldq8, bool write_lo, cudaStream_t st) {
  int64_t blocks = ceil_div(n_clusters * n_slices, CA_THREADS / 32);
  if (blocks > 148 * 8 * 8) blocks = 148 * 8 * 8;
#define GNNLM_CL(W)                                                                                                       \
  cluster_attn_kernel<T, OutT, C, W, GROUP><<<(unsigned)blocks, CA_THREADS, 0, st>>>(                                     \
      (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, node_base, valid_base, cluster_nl, n_clusters, group, n_slices, \
      (OutT*)out, ldo, lo_off, q8, ldq8, write_lo)
  if (centre_only) GNNLM_CL(0);
  else if (wmax <= 1) GNNLM_CL(1);
  else if (wmax <= 3) GNNLM_CL(3);
  else if (wmax <= 5 && !window_forced()) GNNLM_CL(5);      // c = 2 (or c = 3 pruned to reach 2): 15 rows in flight per warp
  else if (wmax <= 7 && !window_forced()) GNNLM_CL(7);
  else
    cluster_attn_window_kernel<T, OutT, C, GROUP><<<(unsigned)blocks, CA_THREADS, 0, st>>>(
        (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, node_base, cluster_nl, n_clusters, group, n_slices, (OutT*)out,
        ldo, lo_off, q8, ldq8, write_lo);
#undef GNNLM_CL
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_cluster_attn");
  return 0;
}

template <typename T, typename OutT, int C>
static int32_t dispatch_group(int wmax, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              const int32_t* node_base, const int32_t* valid_base, const int32_t* cluster_nl,
                              int64_t n_clusters, int centre_only, int group, int n_slices, void* out, int64_t ldo, int lo_off,
                              uint8_t* q8, int64_t ldq8, bool write_lo, cudaStream_t st) {
#define GNNLM_DG(G)                                                                                                          \
  return dispatch_w<T, OutT, C, G>(wmax, q, ldq, k, ldk, v, ldv, node_base, valid_base, cluster_nl, n_clusters, centre_only, \
                                   group, n_slices, out, ldo, lo_off, q8, ldq8, write_lo, st)
  if (group == 32) GNNLM_DG(32);
  if (group == 16) GNNLM_DG(16);
  GNNLM_DG(0);
#undef GNNLM_DG
}

template <int GROUP>
static int32_t launch_hq(int wmax, const __half* q, int64_t ldq, const __half* k, int64_t ldk, const __half* v, int64_t ldv, int64_t qb,
                         int64_t kb, int64_t vb, const int32_t* node_base, const int32_t* valid_base, const int32_t* cluster_nl,
                         int64_t n_clusters, int centre_only, int group, int n_slices, __half* out, int64_t ldo, int lo_off,
                         uint8_t* q8, int64_t ldq8, bool write_lo, const float2* kv_stats, const float* k_c, const float* k_b,
                         const float* v_c, const float* v_b, cudaStream_t st) {
  int64_t blocks = ceil_div(n_clusters * n_slices, CA_THREADS / 32);
  if (blocks > 148 * 8 * 8) blocks = 148 * 8 * 8;
#define GNNLM_CLH(W)                                                                                                               \
  cluster_attn_kernel<__half, __half, 4, W, GROUP, true><<<(unsigned)blocks, CA_THREADS, 0, st>>>(                                 \
      q, ldq, k, ldk, v, ldv, node_base, valid_base, cluster_nl, n_clusters, group, n_slices, out, ldo, lo_off, q8, ldq8, write_lo, \
      qb, kb, vb, kv_stats, k_c, k_b, v_c, v_b)
  if (centre_only) GNNLM_CLH(0);
  else if (wmax <= 1) GNNLM_CLH(1);
  else if (wmax <= 3) GNNLM_CLH(3);
  else if (wmax <= 5) GNNLM_CLH(5);
  else GNNLM_CLH(7);
#undef GNNLM_CLH
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_cluster_attn_hq");
  return 0;
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_hgt_cluster_attn_hq(const void* q, const void* q_lo8, int64_t ldq, const void* k, const void* k_lo8, int64_t ldk,
                                             const void* v, const void* v_lo8, int64_t ldv, const int32_t* node_base,
                                             const int32_t* valid_base, const int32_t* cluster_nl, int64_t n_clusters,
                                             int32_t max_cluster, int32_t centre_only, int32_t H, int32_t d_k, void* out, int64_t ldo,
                                             void* q8v, int64_t ldq8, int32_t write_lo, const float* kv_stats, const float* k_c,
                                             const float* k_b, const float* v_c, const float* v_b, gnnlm_stream_t stream) {
  uint8_t* q8 = reinterpret_cast<uint8_t*>(q8v);
  GNNLM_CHECK_ARG(q && k && v && q_lo8 && k_lo8 && v_lo8 && node_base && cluster_nl && out, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_hq: null pointer");
  GNNLM_CHECK_ARG(!centre_only || valid_base, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_hq: centre_only needs valid_base");
  GNNLM_CHECK_ARG(write_lo || q8, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_hq: write_lo = 0 (hi + companion only) needs q8");
  GNNLM_CHECK_ARG(H > 0 && d_k > 0 && max_cluster >= 1 && max_cluster <= 7, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_cluster_attn_hq: clusters of at most 7 nodes");
  GNNLM_CHECK_ARG(!kv_stats || (centre_only && k_c && k_b && v_c && v_b && (uintptr_t)kv_stats % 8 == 0 && (uintptr_t)k_c % 16 == 0 &&
                                (uintptr_t)k_b % 16 == 0 && (uintptr_t)v_c % 16 == 0 && (uintptr_t)v_b % 16 == 0),
                  GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_hq: the deferred-LayerNorm affine needs the centre-only form and its four vectors");
  const int64_t d = (int64_t)H * d_k;
  GNNLM_CHECK_ARG(d % 128 == 0 && d_k % 4 == 0 && d_k / 4 <= 32 && 32 % (d_k / 4) == 0, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_cluster_attn_hq: unsupported (H=%d, d_k=%d)", H, d_k);
  GNNLM_CHECK_ARG(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 8 == 0 && (uintptr_t)q % 8 == 0 && (uintptr_t)k % 8 == 0 &&
                      (uintptr_t)v % 8 == 0 && (uintptr_t)q_lo8 % 4 == 0 && (uintptr_t)k_lo8 % 4 == 0 && (uintptr_t)v_lo8 % 4 == 0,
                  GNNLM_E_SHAPE, "gnnlm_hgt_cluster_attn_hq: rows must keep 8 B (16-bit plane) / 4 B (byte plane) alignment");
  GNNLM_CHECK_ARG(ldo >= (write_lo ? 2 : 1) * d && (!q8 || ((uintptr_t)q8 % 4 == 0 && ldq8 % 4 == 0 && ldq8 >= 2 * d)), GNNLM_E_SHAPE,
                  "gnnlm_hgt_cluster_attn_hq: split output needs ldo >= 2d (d when only the hi half is written), ldq8 >= 2d");
  if (n_clusters == 0) return 0;
  const int group = d_k / 4, n_slices = (int)(d / 128);
  auto bias = [](const void* hi, const void* lo8) { return (int64_t)((uintptr_t)lo8) - (int64_t)((uintptr_t)hi >> 1); };
  const int64_t qb = bias(q, q_lo8), kb = bias(k, k_lo8), vb = bias(v, v_lo8);
  cudaStream_t st = (cudaStream_t)stream;
#define GNNLM_HQ(G)                                                                                                                \
  return launch_hq<G>(max_cluster, (const __half*)q, ldq, (const __half*)k, ldk, (const __half*)v, ldv, qb, kb, vb, node_base,     \
                      valid_base, cluster_nl, n_clusters, centre_only, group, n_slices, (__half*)out, ldo, (int)d, q8, ldq8,       \
                      write_lo != 0, reinterpret_cast<const float2*>(kv_stats), k_c, k_b, v_c, v_b, st)
  if (group == 32) GNNLM_HQ(32);
  if (group == 16) GNNLM_HQ(16);
  GNNLM_HQ(0);
#undef GNNLM_HQ
}

extern "C" int32_t gnnlm_hgt_cluster_attn(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                          int32_t dtype, const int32_t* node_base, const int32_t* valid_base,
                                          const int32_t* cluster_nl, int64_t n_clusters, int32_t max_cluster,
                                          int32_t centre_only, int32_t H, int32_t d_k, void* out, int32_t out_dtype,
                                          int64_t ldo, gnnlm_stream_t stream) {
  return gnnlm_hgt_cluster_attn_q8(q, ldq, k, ldk, v, ldv, dtype, node_base, valid_base, cluster_nl, n_clusters, max_cluster,
                                   centre_only, H, d_k, out, out_dtype, ldo, nullptr, 0, 1, stream);
}

extern "C" int32_t gnnlm_hgt_cluster_attn_q8(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                             int32_t dtype, const int32_t* node_base, const int32_t* valid_base,
                                             const int32_t* cluster_nl, int64_t n_clusters, int32_t max_cluster,
                                             int32_t centre_only, int32_t H, int32_t d_k, void* out, int32_t out_dtype,
                                             int64_t ldo, void* q8v, int64_t ldq8, int32_t write_lo,
                                             gnnlm_stream_t stream) {
  uint8_t* q8 = reinterpret_cast<uint8_t*>(q8v);
  GNNLM_CHECK_ARG(!q8 || (out_dtype == GNNLM_F16X2 && (uintptr_t)q8 % 4 == 0 && ldq8 % 4 == 0 && ldq8 >= 2 * (int64_t)H * d_k),
                  GNNLM_E_UNSUPPORTED, "gnnlm_hgt_cluster_attn_q8: the e4m3 companion needs a split-fp16 output and ldq8 >= 2d");
  GNNLM_CHECK_ARG(write_lo || q8, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn_q8: write_lo = 0 (hi + companion only) needs q8");
  GNNLM_CHECK_ARG(q && k && v && node_base && cluster_nl && out, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn: null pointer");
  GNNLM_CHECK_ARG(!centre_only || valid_base, GNNLM_E_ARG, "gnnlm_hgt_cluster_attn: centre_only needs valid_base");
  GNNLM_CHECK_ARG(dtype == GNNLM_F32 || dtype == GNNLM_BF16, GNNLM_E_UNSUPPORTED, "gnnlm_hgt_cluster_attn: dtype");
  GNNLM_CHECK_ARG(out_dtype == GNNLM_F32 || out_dtype == GNNLM_BF16 || out_dtype == GNNLM_F16X2, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_cluster_attn: out dtype");
  GNNLM_CHECK_ARG(out_dtype != GNNLM_F16X2 || ldo >= (write_lo ? 2 : 1) * (int64_t)H * d_k, GNNLM_E_SHAPE,
                  "gnnlm_hgt_cluster_attn: split output needs ldo >= 2d (d when only the hi half is written)");
  GNNLM_CHECK_ARG(max_cluster >= 1, GNNLM_E_SHAPE, "gnnlm_hgt_cluster_attn: max_cluster must be >= 1");
  GNNLM_CHECK_ARG(H > 0 && d_k > 0, GNNLM_E_SHAPE, "gnnlm_hgt_cluster_attn: H and d_k must be positive");
  const int64_t d = (int64_t)H * d_k;
  const int Cs = dtype == GNNLM_F32 ? 4 : 8;                 // 16 B per lane per row
  // a warp covers 32*Cs features; a head (d_k features) must live inside one warp and heads may not straddle warps
  GNNLM_CHECK_ARG(d % (32 * Cs) == 0 && d_k % Cs == 0 && d_k / Cs <= 32 && 32 % (d_k / Cs) == 0, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_cluster_attn: unsupported (H=%d, d_k=%d) -- use gnnlm_hgt_edge_attn", H, d_k);
  GNNLM_CHECK_ARG(ldq % Cs == 0 && ldk % Cs == 0 && ldv % Cs == 0 && ldo % 8 == 0, GNNLM_E_SHAPE,
                  "gnnlm_hgt_cluster_attn: leading dimensions must keep 16 B alignment");
  if (n_clusters == 0) return 0;
  const int group = d_k / Cs, n_slices = (int)(d / (32 * Cs));
  cudaStream_t st = (cudaStream_t)stream;
#define GNNLM_DW(T, OT, C)                                                                                                    \
  return dispatch_group<T, OT, C>(max_cluster, q, ldq, k, ldk, v, ldv, node_base, valid_base, cluster_nl, n_clusters, centre_only, \
                              group, n_slices, out, ldo, (int)d, q8, ldq8, write_lo != 0, st)
  if (dtype == GNNLM_F32 && out_dtype == GNNLM_F32) GNNLM_DW(float, float, 4);
  if (dtype == GNNLM_F32 && out_dtype == GNNLM_F16X2) GNNLM_DW(float, __half, 4);
  if (out_dtype == GNNLM_F16X2) GNNLM_DW(__nv_bfloat16, __half, 8);
  if (dtype == GNNLM_F32) GNNLM_DW(float, __nv_bfloat16, 4);
  if (out_dtype == GNNLM_F32) GNNLM_DW(__nv_bfloat16, float, 8);
  GNNLM_DW(__nv_bfloat16, __nv_bfloat16, 8);
#undef GNNLM_DW
}
