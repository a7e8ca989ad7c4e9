// HGT edge attention: per-edge-type Q.K' score, segmented softmax by destination, weighted V' sum,
// cross-edge-type mean -- one warp per destination, shuffle reductions, no atomics.
//
// Replaces, per canonical edge type, the DGL sequence of HGTLayer.forward
// (reference: fairseq/models/hgt.py:350-358,383-386): apply_edges(fn.v_dot_u('q','k','t')) (SDDMM),
// `* relation_pri / sqrt_dk`, dgl.ops.edge_softmax(norm_by='dst'), and
// multi_update_all({etype: (u_mul_e, sum)}, cross_reducer='mean').  The relation transforms
// (:347-348) and the pri/sqrt(d_k) factor are folded into the K'/V' projection weights by the host
// module, so `k`/`v` here are already K', V'.
//
// DGL semantics restated (DGL itself is absent): softmax is per destination and per
// head, max-subtracted; a destination without in-edges receives 0 from that edge type; the
// cross-type mean divides by the number of edge types targeting the node type regardless of degree
// (out_scale = 1/n_etypes, accumulate chains the edge types).
//
// Layout: q [n_dst, d], k/v [n_src, d] row-major, d = H*d_k.  A warp owns one destination; lane l
// owns the C = d/32 contiguous features [l*C, (l+1)*C), all inside head l / (32/H).  Scores are
// reduced over the 32/H lanes of a head with xor-shuffles; softmax is online (running max / sum),
// so each K'/V' row is read exactly once per in-edge and nothing intermediate ([E,H] scores, [E,H,d_k]
// messages) is written.  Algorithmic HBM bytes per layer and edge type:
// N_src*2*d*s + N_dst*d*s (q) + N_dst*d*4 (out) + E*4 + (N_dst+1)*4.
#include "common.cuh"

namespace gnnlm {

constexpr int EA_THREADS = 256;

template <typename T, int C, int UNROLL, int GROUP = 0>      // GROUP: lanes per head when known at compile time (0 = `group`)
__device__ __forceinline__ void attend_range(const float (&q)[C], const T* __restrict__ k, int64_t ldk,
                                             const T* __restrict__ v, int64_t ldv, const int32_t* __restrict__ indices,
                                             int64_t e_begin, int64_t e_end, int col, int group, float& m_run,
                                             float& l_run, float (&acc)[C]) {
  // UNROLL in-edges in flight per pass, predicated (a chain node has <= 3 in-edges: one pass, one
  // round of memory latency).  `col` = first feature owned by this lane.
  for (int64_t e = e_begin; e < e_end; e += UNROLL) {
    float kk[UNROLL][C], vv[UNROLL][C];
    float s[UNROLL];
    int64_t src[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const bool ok = e + u < e_end;
      src[u] = ok ? (indices ? (int64_t)__ldg(indices + e + u) : (e + u)) : -1;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (src[u] >= 0) {
        load_row<T, C>(k + src[u] * ldk + col, kk[u]);
        load_row<T, C>(v + src[u] * ldv + col, vv[u]);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) kk[u][c] = vv[u][c] = 0.f;
      }
    }
    float mx = m_run;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float p = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) p = fmaf(q[c], kk[u][c], p);
      if constexpr (GROUP > 0) {
#pragma unroll
        for (int o = GROUP >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      } else {
        for (int o = group >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      }
      s[u] = src[u] >= 0 ? p : -INFINITY;
      mx = fmaxf(mx, s[u]);
    }
    const float corr = __expf(m_run - mx);          // m_run == -inf on first use -> 0
    l_run *= corr;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] *= corr;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const float w = __expf(s[u] - mx);            // -inf -> 0 for the predicated-off slots
      l_run += w;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(w, vv[u][c], acc[c]);
    }
    m_run = mx;
  }
}

template <int C>
__device__ __forceinline__ void write_out(float* __restrict__ o, const float (&acc)[C], float l_run, float out_scale,
                                          int accumulate) {
  const float inv = l_run > 0.f ? out_scale / l_run : 0.f;     // zero in-degree -> 0 (fn.sum semantics)
  float r[C];
  if (accumulate) {
    load_f32<C>(o, r);
#pragma unroll
    for (int c = 0; c < C; ++c) r[c] = fmaf(acc[c], inv, r[c]);
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) r[c] = acc[c] * inv;
  }
  store_f32<C>(o, r);
}

// CSR edge attention.  Work item = (destination, feature slice of 32*C floats): a warp reads 128*C
// contiguous bytes of each Q / K' / V' row, so registers stay small (many warps resident, several rows in
// flight per warp) and every row segment is one coalesced request.  `group` = lanes per head.
template <typename T, int C, int UNROLL, int GROUP>
__global__ void __launch_bounds__(EA_THREADS) edge_attn_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k,
                                                               int64_t ldk, const T* __restrict__ v, int64_t ldv,
                                                               const int32_t* __restrict__ indptr,
                                                               const int32_t* __restrict__ indices,
                                                               const int32_t* __restrict__ dst_ids, int64_t n_dst_cap,
                                                               const int32_t* __restrict__ n_dst_dev, int group, int n_slices,
                                                               float* __restrict__ out, int64_t ldo, float out_scale,
                                                               int accumulate) {
  const int64_t n_items = live_rows(n_dst_cap, n_dst_dev) * n_slices;
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += warps) {
    const int64_t i = it / n_slices;
    const int col = ((int)(it % n_slices) * 32 + lane) * C;
    const int64_t row = dst_ids ? (int64_t)__ldg(dst_ids + i) : i;
    const int64_t e0 = __ldg(indptr + row), e1 = __ldg(indptr + row + 1);
    float qr[C], acc[C];
    load_row<T, C>(q + i * ldq + col, qr);
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    attend_range<T, C, UNROLL, GROUP>(qr, k, ldk, v, ldv, indices, e0, e1, col, group, m_run, l_run, acc);
    write_out<C>(out + i * ldo + col, acc, l_run, out_scale, accumulate);
  }
}

// Implicit causal attention inside blocks of L tokens: destination g = b*L + t attends to sources
// b*L + [max(0, t-ctx+1), t].  v1: warp per destination (K'/V' rows come from L2: a 3072-token block is
// 2 x 12.6 MB in fp32).
template <typename T, int C, int UNROLL>
__global__ void __launch_bounds__(EA_THREADS) causal_attn_kernel(const T* __restrict__ q, int64_t ldq,
                                                                 const T* __restrict__ k, int64_t ldk,
                                                                 const T* __restrict__ v, int64_t ldv, int64_t B,
                                                                 int64_t L, int64_t ctx, int group, float* __restrict__ out,
                                                                 int64_t ldo, float out_scale, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t n = B * L;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += warps) {
    // longest rows first so the tail of the grid is made of short rows
    const int64_t g = n - 1 - w;
    const int64_t b = g / L, t = g % L;
    const int64_t lo = (ctx > 0 && t + 1 > ctx) ? t + 1 - ctx : 0;
    float qr[C], acc[C];
    load_row<T, C>(q + g * ldq + lane * C, qr);
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    attend_range<T, C, UNROLL>(qr, k, ldk, v, ldv, nullptr, b * L + lo, b * L + t + 1, lane * C, group, m_run, l_run, acc);
    write_out<C>(out + g * ldo + lane * C, acc, l_run, out_scale, accumulate);
  }
}

// ---------------------------------------------------------------------------------------------------
// Tiled ("flash") form of the implicit causal attention for fp32 and d_k in {64, 128}: one CTA owns a
// 64-query tile of one head, streams 32-key K'/V' tiles through shared memory, keeps the running
// (max, sum) per query row and the 64 x d_k output tile in registers.  K'/V' of a 3072-token block are
// re-read L/64 times from L2 instead of L times (warp-per-destination form above), and all arithmetic
// is fp32 FMA (exact-parity modes).  Thread layout keeps every shared-memory request to one wavefront:
// a warp covers 8 query-row groups x 4 column groups, rows padded by 16 B.
constexpr int FA_BQ = 64, FA_BK = 32, FA_THREADS = 256;

template <int DK>
__global__ void __launch_bounds__(FA_THREADS, 2)
    causal_flash_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                        const float* __restrict__ v, int64_t ldv, int64_t L, int64_t ctx, float* __restrict__ out,
                        int64_t ldo, float out_scale, int accumulate) {
  constexpr int LD = DK + 4;                 // padded row (floats)
  constexpr int LDP = FA_BK + 4;
  constexpr int OC = DK / 16;                // output columns per thread (8 for d_k=128)
  extern __shared__ __align__(16) float fsm[];
  float* Qs = fsm;                           // [64][LD]
  float* Ks = Qs + FA_BQ * LD;               // [32][LD]
  float* Vs = Ks + FA_BK * LD;               // [32][LD]
  float* Ps = Vs + FA_BK * LD;               // [64][LDP]
  float* alpha_s = Ps + FA_BQ * LDP;         // [64]

  const int n_qt = (int)((L + FA_BQ - 1) / FA_BQ);
  const int qt = n_qt - 1 - (int)blockIdx.x;               // longest tiles first
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t row0 = b * L;                              // first token of this block
  const int q0 = qt * FA_BQ;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int ty = (w & 1) * 8 + (lane >> 2);                // 0..15
  const int tx = (w >> 1) * 4 + (lane & 3);                // 0..15
  const int64_t col0 = (int64_t)h * DK;

  // Q tile -> smem (rows beyond L are zero)
  for (int i = tid; i < FA_BQ * (DK / 4); i += FA_THREADS) {
    const int r = i / (DK / 4), c4 = i % (DK / 4);
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < L) t = __ldg(reinterpret_cast<const float4*>(q + (row0 + q0 + r) * ldq + col0) + c4);
    *reinterpret_cast<float4*>(Qs + r * LD + c4 * 4) = t;
  }
  float o[4][OC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < OC; ++c) o[i][c] = 0.f;
  // softmax bookkeeping: thread (tid/4) owns row tid/4 together with 3 neighbours (8 columns each)
  const int srow = tid >> 2, spart = tid & 3;
  float m_run = -INFINITY, l_run = 0.f;

  const int q_hi = min(q0 + FA_BQ, (int)L) - 1;            // last query of the tile
  int kt_lo = 0;
  if (ctx > 0) {
    const int64_t first_key = (int64_t)q0 + 1 - ctx;       // earliest key any query of the tile may see
    if (first_key > 0) kt_lo = (int)(first_key / FA_BK);
  }
  const int kt_hi = q_hi / FA_BK;                          // inclusive
  for (int kt = kt_lo; kt <= kt_hi; ++kt) {
    const int k0 = kt * FA_BK;
    __syncthreads();                                       // previous tile's P.V done (and Q visible on first pass)
    for (int i = tid; i < FA_BK * (DK / 4); i += FA_THREADS) {
      const int r = i / (DK / 4), c4 = i % (DK / 4);
      float4 tk = make_float4(0.f, 0.f, 0.f, 0.f), tv = tk;
      if (k0 + r < L) {
        tk = __ldg(reinterpret_cast<const float4*>(k + (row0 + k0 + r) * ldk + col0) + c4);
        tv = __ldg(reinterpret_cast<const float4*>(v + (row0 + k0 + r) * ldv + col0) + c4);
      }
      *reinterpret_cast<float4*>(Ks + r * LD + c4 * 4) = tk;
      *reinterpret_cast<float4*>(Vs + r * LD + c4 * 4) = tv;
    }
    __syncthreads();
    // ---- S = Q K^T: thread -> rows ty*4+i, key columns tx*2+j
    float s[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = 0.f;
#pragma unroll 8
    for (int d4 = 0; d4 < DK / 4; ++d4) {
      float4 a[4], bb[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(Qs + (ty * 4 + i) * LD + d4 * 4);
#pragma unroll
      for (int j = 0; j < 2; ++j) bb[j] = *reinterpret_cast<const float4*>(Ks + (tx * 2 + j) * LD + d4 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          s[i][j] = fmaf(a[i].x, bb[j].x, s[i][j]);
          s[i][j] = fmaf(a[i].y, bb[j].y, s[i][j]);
          s[i][j] = fmaf(a[i].z, bb[j].z, s[i][j]);
          s[i][j] = fmaf(a[i].w, bb[j].w, s[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int qi = q0 + ty * 4 + i, kj = k0 + tx * 2 + j;
        const bool ok = kj <= qi && kj < L && (ctx <= 0 || qi - kj < ctx);
        Ps[(ty * 4 + i) * LDP + tx * 2 + j] = ok ? s[i][j] : -INFINITY;
      }
    __syncthreads();
    // ---- online softmax over the 32 new columns of each row
    {
      float* pr = Ps + srow * LDP + spart * 8;
      float4 x0 = *reinterpret_cast<float4*>(pr), x1 = *reinterpret_cast<float4*>(pr + 4);
      float mx = fmaxf(fmaxf(fmaxf(x0.x, x0.y), fmaxf(x0.z, x0.w)), fmaxf(fmaxf(x1.x, x1.y), fmaxf(x1.z, x1.w)));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);
      const float base = m_new == -INFINITY ? 0.f : m_new;            // fully masked so far -> all p = 0
      x0.x = __expf(x0.x - base); x0.y = __expf(x0.y - base); x0.z = __expf(x0.z - base); x0.w = __expf(x0.w - base);
      x1.x = __expf(x1.x - base); x1.y = __expf(x1.y - base); x1.z = __expf(x1.z - base); x1.w = __expf(x1.w - base);
      *reinterpret_cast<float4*>(pr) = x0;
      *reinterpret_cast<float4*>(pr + 4) = x1;
      float sum = (x0.x + x0.y) + (x0.z + x0.w) + (x1.x + x1.y) + (x1.z + x1.w);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float al = m_run == -INFINITY ? 0.f : __expf(m_run - base);
      l_run = l_run * al + sum;
      m_run = m_new;
      if (spart == 0) alpha_s[srow] = al;
    }
    __syncthreads();
    // ---- O = O * alpha + P V: thread -> rows ty*4+i, dims tx*OC..
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float al = alpha_s[ty * 4 + i];
#pragma unroll
      for (int c = 0; c < OC; ++c) o[i][c] *= al;
    }
#pragma unroll 2
    for (int j4 = 0; j4 < FA_BK / 4; ++j4) {
      float4 p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = *reinterpret_cast<const float4*>(Ps + (ty * 4 + i) * LDP + j4 * 4);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float vv[OC];
#pragma unroll
        for (int c4 = 0; c4 < OC / 4; ++c4) {
          const float4 t = *reinterpret_cast<const float4*>(Vs + (j4 * 4 + jj) * LD + tx * OC + c4 * 4);
          vv[c4 * 4] = t.x; vv[c4 * 4 + 1] = t.y; vv[c4 * 4 + 2] = t.z; vv[c4 * 4 + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float pij = jj == 0 ? p[i].x : (jj == 1 ? p[i].y : (jj == 2 ? p[i].z : p[i].w));
#pragma unroll
          for (int c = 0; c < OC; ++c) o[i][c] = fmaf(pij, vv[c], o[i][c]);
        }
      }
    }
  }
  __syncthreads();
  if (spart == 0) alpha_s[srow] = l_run;                   // reuse as the row-sum table
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi >= L) continue;
    const float l = alpha_s[ty * 4 + i];
    const float inv = l > 0.f ? out_scale / l : 0.f;
    float* dst = out + (row0 + qi) * ldo + col0 + tx * OC;
#pragma unroll
    for (int c4 = 0; c4 < OC / 4; ++c4) {
      float4 r = make_float4(o[i][c4 * 4] * inv, o[i][c4 * 4 + 1] * inv, o[i][c4 * 4 + 2] * inv, o[i][c4 * 4 + 3] * inv);
      if (accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(dst + c4 * 4);
        r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
      }
      *reinterpret_cast<float4*>(dst + c4 * 4) = r;
    }
  }
}

template <int DK>
static int32_t launch_flash(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, int64_t B,
                            int64_t L, int64_t ctx, int H, float* out, int64_t ldo, float out_scale, int accumulate,
                            cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)(FA_BQ + 2 * FA_BK) * (DK + 4) + (size_t)FA_BQ * (FA_BK + 4) + FA_BQ);
  static bool attr = false;
  if (!attr) {
    GNNLM_CUDA(cudaFuncSetAttribute(causal_flash_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((unsigned)ceil_div(L, FA_BQ), (unsigned)H, (unsigned)B);
  causal_flash_kernel<DK><<<grid, FA_THREADS, smem, st>>>(q, ldq, k, ldk, v, ldv, L, ctx, out, ldo, out_scale, accumulate);
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_causal_attn(flash)");
  return 0;
}

template <typename T, int C>
static int32_t launch_edge(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           const int32_t* indptr, const int32_t* indices, const int32_t* dst_ids, int64_t n_dst_cap,
                           const int32_t* n_dst_dev, int group, int n_slices, float* out, int64_t ldo, float out_scale,
                           int accumulate, cudaStream_t st) {
  constexpr int UNROLL = 4;
  const int64_t items = n_dst_cap * n_slices;
  int64_t blocks = ceil_div(items, EA_THREADS / 32);
  const int64_t max_blocks = 148 * 8 * 8;
  if (blocks > max_blocks) blocks = max_blocks;
#define GNNLM_EA(G)                                                                                                         \
  edge_attn_kernel<T, C, UNROLL, G><<<(unsigned)blocks, EA_THREADS, 0, st>>>(                                                \
      (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, indptr, indices, dst_ids, n_dst_cap, n_dst_dev, group, n_slices, \
      out, ldo, out_scale, accumulate)
  if (group == 32) GNNLM_EA(32);           // d_k = 128 fp32 / 256 bf16: fully unrolled reductions, no divergent-collective code
  else if (group == 16) GNNLM_EA(16);
  else GNNLM_EA(0);
#undef GNNLM_EA
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_edge_attn");
  return 0;
}

template <typename T, int C>
static int32_t launch_causal(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, int64_t B,
                             int64_t L, int64_t ctx, int group, float* out, int64_t ldo, float out_scale, int accumulate,
                             cudaStream_t st) {
  constexpr int UNROLL = C >= 32 ? 2 : 4;
  int64_t blocks = ceil_div(B * L, EA_THREADS / 32);
  causal_attn_kernel<T, C, UNROLL><<<(unsigned)blocks, EA_THREADS, 0, st>>>((const T*)q, ldq, (const T*)k, ldk,
                                                                            (const T*)v, ldv, B, L, ctx, group, out, ldo,
                                                                            out_scale, accumulate);
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_causal_attn");
  return 0;
}

// Two lane mappings: "whole" -- a warp spans the whole row, d/32 features per lane, 32/H lanes per head (H a power of
// two); "sliced" -- a lane owns 16 B, a warp spans 32 lanes * 16 B of the row and the row takes d / that many slices
// (any H as long as a head's lanes divide the warp: d_k = 64, 128, ... in fp32).
static int32_t check_shape(const char* who, int32_t H, int32_t d_k, int32_t dtype, int64_t ldq, int64_t ldk, int64_t ldv,
                           int64_t ldo, int* C_out, int* group_out, int* slices_out, bool need_whole) {
  GNNLM_CHECK_ARG(dtype == GNNLM_F32 || dtype == GNNLM_BF16, GNNLM_E_UNSUPPORTED, "%s: dtype must be F32 or BF16", who);
  GNNLM_CHECK_ARG(H > 0 && d_k > 0, GNNLM_E_SHAPE, "%s: H and d_k must be positive", who);
  const int64_t d = (int64_t)H * d_k;
  const int Cs = dtype == GNNLM_F32 ? 4 : 8;                   // 16 B per lane per row
  const bool sliced_ok = d % (32 * Cs) == 0 && d_k % Cs == 0 && d_k / Cs <= 32 && 32 % (d_k / Cs) == 0;
  const int64_t Cw = d / 32;
  const bool whole_ok = H <= 32 && (H & (H - 1)) == 0 && d % 32 == 0 &&
                        (Cw == 1 || Cw == 2 || Cw == 4 || Cw == 8 || Cw == 16 || Cw == 32);
  GNNLM_CHECK_ARG(whole_ok || (sliced_ok && !need_whole), GNNLM_E_UNSUPPORTED,
                  "%s: unsupported head layout (H=%d, d_k=%d): need H a power of two <= 32 with H*d_k/32 in {1..32}%s", who, H,
                  d_k, need_whole ? "" : ", or d_k a multiple of 4 (fp32) / 8 (bf16) whose lanes divide a warp");
  if (!need_whole && sliced_ok && (Cw > Cs || !whole_ok)) {
    *C_out = Cs;
    *group_out = d_k / Cs;                                     // lanes per head inside a slice
    *slices_out = (int)(d / (32 * Cs));
  } else {
    *C_out = (int)Cw;
    *group_out = 32 / H;
    *slices_out = 1;
  }
  GNNLM_CHECK_ARG(ldq % *C_out == 0 && ldk % *C_out == 0 && ldv % *C_out == 0 && ldo % *C_out == 0, GNNLM_E_SHAPE,
                  "%s: leading dimensions must be multiples of the per-lane width (%d)", who, *C_out);
  return 0;
}

#define DISPATCH_C(FN, T, ...)                   \
  switch (C) {                                   \
    case 1: return FN<T, 1>(__VA_ARGS__);        \
    case 2: return FN<T, 2>(__VA_ARGS__);        \
    case 4: return FN<T, 4>(__VA_ARGS__);        \
    case 8: return FN<T, 8>(__VA_ARGS__);        \
    case 16: return FN<T, 16>(__VA_ARGS__);      \
    default: return FN<T, 32>(__VA_ARGS__);      \
  }

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_hgt_edge_attn(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                       int32_t dtype, const int32_t* indptr, const int32_t* indices,
                                       const int32_t* dst_ids, int64_t n_dst_cap, const int32_t* n_dst_dev, int32_t H,
                                       int32_t d_k, float* out, int64_t ldo, float out_scale, int32_t accumulate,
                                       gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(q && k && v && indptr && out, GNNLM_E_ARG, "gnnlm_hgt_edge_attn: null pointer");
  int C, group, n_slices;
  int32_t rc = check_shape("gnnlm_hgt_edge_attn", H, d_k, dtype, ldq, ldk, ldv, ldo, &C, &group, &n_slices, false);
  if (rc) return rc;
  if (n_dst_cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GNNLM_F32) {
    DISPATCH_C(launch_edge, float, q, ldq, k, ldk, v, ldv, indptr, indices, dst_ids, n_dst_cap, n_dst_dev, group, n_slices,
               out, ldo, out_scale, accumulate, st)
  } else {
    DISPATCH_C(launch_edge, __nv_bfloat16, q, ldq, k, ldk, v, ldv, indptr, indices, dst_ids, n_dst_cap, n_dst_dev, group,
               n_slices, out, ldo, out_scale, accumulate, st)
  }
}

extern "C" int32_t gnnlm_hgt_causal_attn(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                         int64_t ldv, int32_t dtype, int64_t B, int64_t L, int64_t intra_ctx, int32_t H,
                                         int32_t d_k, float* out, int64_t ldo, float out_scale, int32_t accumulate,
                                         gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(q && k && v && out, GNNLM_E_ARG, "gnnlm_hgt_causal_attn: null pointer");
  GNNLM_CHECK_ARG(B >= 0 && L > 0, GNNLM_E_SHAPE, "gnnlm_hgt_causal_attn: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const bool al16 = ((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                    ((uintptr_t)out % 16 == 0) && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0;
  if (dtype == GNNLM_F32 && al16 && L >= 64 && B <= 65535 && B > 0 && H > 0 && (d_k == 128 || d_k == 64)) {   // any H
    if (d_k == 128)
      return launch_flash<128>((const float*)q, ldq, (const float*)k, ldk, (const float*)v, ldv, B, L, intra_ctx, H, out,
                               ldo, out_scale, accumulate, st);
    return launch_flash<64>((const float*)q, ldq, (const float*)k, ldk, (const float*)v, ldv, B, L, intra_ctx, H, out, ldo,
                            out_scale, accumulate, st);
  }
  int C, group, n_slices;
  int32_t rc = check_shape("gnnlm_hgt_causal_attn", H, d_k, dtype, ldq, ldk, ldv, ldo, &C, &group, &n_slices, true);
  if (rc) return rc;
  if (B == 0) return 0;
  if (dtype == GNNLM_F32) {
    DISPATCH_C(launch_causal, float, q, ldq, k, ldk, v, ldv, B, L, intra_ctx, group, out, ldo, out_scale, accumulate, st)
  } else {
    DISPATCH_C(launch_causal, __nv_bfloat16, q, ldq, k, ldk, v, ldv, B, L, intra_ctx, group, out, ldo, out_scale,
               accumulate, st)
  }
}
