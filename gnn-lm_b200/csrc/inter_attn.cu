// ('ntgt','inter','tgt') attention without projecting the centre nodes.
//
// HGTLayer.forward (reference: fairseq/models/hgt.py:339-358,383-386) computes, for the inter edge type,
//   K'[c] = h_c W_k'^T + b_k',  V'[c] = h_c W_v'^T + b_v'     for every centre node c   (2 d^2 MACs per centre, ~97 k centres)
//   s[t,c,h] = <q[t,h], K'[c,h]>,   alpha = softmax_c(s),   out[t,h] = sum_c alpha[t,c,h] V'[c,h]
// (relation_att / relation_msg / relation_pri / sqrt(d_k) already folded into W_k', W_v').  In the graphs of
// new_build_graph every centre node feeds exactly ONE target token (token_block_dataset.py:372-373), so the projections
// can move to the token side, where there are k = 32 times fewer rows:
//   s[t,c,h]  = <h_c, q~[t,h]> + <q[t,h], b_k'[h]>,   q~[t,h] = W_k'[h]^T q[t,h]  in R^d      (the bias term is the same for every
//                                                                                              c of a token: softmax drops it)
//   out[t,h]  = W_v'[h] a[t,h] + b_v'[h] * [deg(t) > 0],   a[t,h] = sum_c alpha[t,c,h] h_c  in R^d
// i.e. two per-head GEMMs over the T target rows (gemm_tcgen05.cu, batched) around the kernel of this file, which per token streams
// its <= k centre rows once, scores them against the H transformed queries, and writes the H alpha-weighted row sums:
// d reads + H*d writes per token instead of 2 d^2 MACs per centre.  Same result up to fp32 re-association.
//
// Three forms, chosen by the launcher:
//   inter_mma_kernel    split-fp16 or bf16 centre rows, H = 8, d in {512, 768, 1024}: both contractions on mma.sync tensor cores,
//                       persistent CTA per SM, cp.async tile ring (the default path of MATH_F16X3 / MATH_BF16)
//   inter_regq_kernel   fp32 rows, or H = 4: CUDA cores, q~ in registers, persistent, cp.async tile ring
//   inter_fused_kernel  H = 12 / 16 (or GNNLM_INTER_KERNEL=smemq): CUDA cores, q~ in shared memory, one CTA (256 threads) per token,
//                       two CTAs per SM; centre rows in tiles of 16 through shared memory (fp32, row stride d + 4 floats: conflict-free
//                       for both access patterns), online softmax across tiles.  Phase 1: warp -> head, lane -> (half of the
//                       columns, row).  Phase 2: thread -> 4 columns of all H sums.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "mma_sync.cuh"

namespace gnnlm {

constexpr int IA_THREADS = 256;
constexpr int IA_ROWS = 16;          // centre rows per tile: 2 CTAs per SM at d = 1024, H = 8 (copy of one overlaps the math of the other)

template <typename InT>
__device__ __forceinline__ float4 ia_load4(const InT* row, int64_t c, int64_t d);
template <>
__device__ __forceinline__ float4 ia_load4<float>(const float* row, int64_t c, int64_t) {
  return __ldg(reinterpret_cast<const float4*>(row + c));
}
template <>
__device__ __forceinline__ float4 ia_load4<__half>(const __half* row, int64_t c, int64_t d) {     // split fp16: hi | lo
  return join4_f16(__ldg(reinterpret_cast<const uint2*>(row + c)), __ldg(reinterpret_cast<const uint2*>(row + d + c)));
}
template <>
__device__ __forceinline__ float4 ia_load4<__nv_bfloat16>(const __nv_bfloat16* row, int64_t c, int64_t) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(row + c));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename InT, int H>
__global__ void __launch_bounds__(IA_THREADS, 2) inter_fused_kernel(const float* __restrict__ qt, int64_t q_hs,   // [H, T, d], head stride
                                                                 const InT* __restrict__ hc, int64_t ldh,
                                                                 const int32_t* __restrict__ indptr, int64_t t0, int64_t d,
                                                                 __half* __restrict__ a_out, int64_t a_hs, int64_t lda,
                                                                 const float* __restrict__ bias_v, float out_scale,
                                                                 float* __restrict__ t_agg, int64_t ldt) {
  constexpr int HH = H;                                    // heads per thread in phase 2
  extern __shared__ __align__(16) float ia_smem[];
  const int64_t ldr = d + 4;
  float* rows = ia_smem;                                   // [IA_ROWS][d + 4]
  float* qs = rows + IA_ROWS * ldr;                        // [H][d]
  float* ss = qs + (int64_t)H * d;                         // [IA_ROWS][H] scores, then softmax numerators, of the current tile
  float* corr_s = ss + IA_ROWS * H;                        // [H] rescale of the running sums
  float* l_s = corr_s + H;                                 // [H] running denominators
  float* m_s = l_s + H;                                    // [H] running maxima

  const int64_t t = t0 + blockIdx.x;                       // token (row of qt / a_out / t_agg)
  const int64_t tl = blockIdx.x;                           // token inside this indptr (chunk graphs start at 0)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t e0 = __ldg(indptr + tl), e1 = __ldg(indptr + tl + 1);
  const int64_t deg = e1 - e0;

  const int di = (int)d;                                    // 32-bit index arithmetic below (d <= 1024)
  for (int h = warp; h < H; h += IA_THREADS / 32) {         // warp per head row: no divisions in the copy loops
    const float* src = qt + h * q_hs + t * d;
    for (int c = lane * 4; c < di; c += 128)
      *reinterpret_cast<float4*>(qs + h * di + c) = __ldg(reinterpret_cast<const float4*>(src + c));
  }
  if (tid < H) {
    l_s[tid] = 0.f;
    m_s[tid] = -INFINITY;
  }
  // phase 2 ownership: 4 columns x HH heads per thread
  const int col = tid * 4;
  constexpr int h0 = 0;
  float acc[HH][4];
#pragma unroll
  for (int h = 0; h < HH; ++h) acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;
  // phase 1 ownership: warp -> head (mod 8); lane -> (half of the columns, row)
  const int p1_r = lane & 15, p1_ch = lane >> 4;
  const int half_q = (int)(d / 8);                         // float4 per column half

  for (int64_t r0 = 0; r0 < deg; r0 += IA_ROWS) {
    const int nr = (int)((deg - r0) < IA_ROWS ? (deg - r0) : IA_ROWS);
    __syncthreads();                                       // previous tile fully consumed (and qs / l_s / m_s visible)
    // global -> shared, one warp per centre row (two rows per warp per tile), all loads of a row issued before its stores
    for (int r = warp; r < nr; r += IA_THREADS / 32) {
      const InT* src = hc + (e0 + r0 + r) * ldh;
      float* dst = rows + r * (int)ldr;
      float4 buf[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = lane * 4 + u * 128;
        if (c < di) buf[u] = ia_load4<InT>(src, c, d);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = lane * 4 + u * 128;
        if (c < di) *reinterpret_cast<float4*>(dst + c) = buf[u];
      }
    }
    __syncthreads();
    // ---- phase 1a: scores of the tile
#pragma unroll
    for (int hi = 0; hi < (H + 7) / 8; ++hi) {
      const int h = (warp & 7) + hi * 8;
      if (h < H) {
        float s = 0.f;
        if (p1_r < nr) {
          const float4* rp = reinterpret_cast<const float4*>(rows + p1_r * (int)ldr) + p1_ch * half_q;
          const float4* qp = reinterpret_cast<const float4*>(qs + h * di) + p1_ch * half_q;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f, s6 = 0.f, s7 = 0.f;
          int j = 0;
#pragma unroll 4
          for (; j + 1 < half_q; j += 2) {
            const float4 a = rp[j], b = qp[j], a2 = rp[j + 1], b2 = qp[j + 1];
            s0 = fmaf(a.x, b.x, s0); s1 = fmaf(a.y, b.y, s1); s2 = fmaf(a.z, b.z, s2); s3 = fmaf(a.w, b.w, s3);
            s4 = fmaf(a2.x, b2.x, s4); s5 = fmaf(a2.y, b2.y, s5); s6 = fmaf(a2.z, b2.z, s6); s7 = fmaf(a2.w, b2.w, s7);
          }
          if (j < half_q) {
            const float4 a = rp[j], b = qp[j];
            s0 = fmaf(a.x, b.x, s0); s1 = fmaf(a.y, b.y, s1); s2 = fmaf(a.z, b.z, s2); s3 = fmaf(a.w, b.w, s3);
          }
          s = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
        }
        s += __shfl_xor_sync(0xffffffffu, s, 16);            // the two column halves
        if (p1_ch == 0) ss[p1_r * H + h] = p1_r < nr ? s : -INFINITY;
      }
    }
    __syncthreads();
    // ---- phase 1b: online softmax bookkeeping (warp h, lane -> row)
    for (int h = warp; h < H; h += IA_THREADS / 32) {
      const float s = lane < IA_ROWS ? ss[lane * H + h] : -INFINITY;
      const float m_old = m_s[h];
      const float mx = fmaxf(m_old, warp_max(s));
      const float p = lane < nr ? __expf(s - mx) : 0.f;
      const float tile_sum = warp_sum(p);
      __syncwarp();
      if (lane < IA_ROWS) ss[lane * H + h] = p;
      if (lane == 0) {
        const float c = __expf(m_old - mx);                  // first tile: exp(-inf) = 0
        corr_s[h] = c;
        l_s[h] = l_s[h] * c + tile_sum;
        m_s[h] = mx;
      }
    }
    __syncthreads();
    // ---- phase 2: rescale and add the tile's weighted rows
    if (col < d) {
#pragma unroll
      for (int h = 0; h < HH; ++h) {
        const float cr = corr_s[h0 + h];
        acc[h][0] *= cr; acc[h][1] *= cr; acc[h][2] *= cr; acc[h][3] *= cr;
      }
#pragma unroll 4
      for (int r = 0; r < nr; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(rows + r * (int)ldr + col);
        float pr[HH];                                        // the row's H softmax numerators: H / 4 broadcast 16 B reads
#pragma unroll
        for (int h4 = 0; h4 < HH / 4; ++h4) {
          const float4 p4 = *reinterpret_cast<const float4*>(ss + r * H + h0 + 4 * h4);
          pr[4 * h4] = p4.x; pr[4 * h4 + 1] = p4.y; pr[4 * h4 + 2] = p4.z; pr[4 * h4 + 3] = p4.w;
        }
#pragma unroll
        for (int h = 0; h < HH; ++h) {
          const float p = pr[h];
          acc[h][0] = fmaf(p, x.x, acc[h][0]); acc[h][1] = fmaf(p, x.y, acc[h][1]);
          acc[h][2] = fmaf(p, x.z, acc[h][2]); acc[h][3] = fmaf(p, x.w, acc[h][3]);
        }
      }
    }
  }
  __syncthreads();
  if (col < d) {
#pragma unroll
    for (int h = 0; h < HH; ++h) {
      const float inv = deg > 0 ? 1.f / l_s[h0 + h] : 0.f;
      uint2 hi, lo;
      split4_f16(acc[h][0] * inv, acc[h][1] * inv, acc[h][2] * inv, acc[h][3] * inv, hi, lo);
      __half* o = a_out + (h0 + h) * a_hs + t * lda + col;
      *reinterpret_cast<uint2*>(o) = hi;
      *reinterpret_cast<uint2*>(o + d) = lo;
    }
    {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (deg > 0 && bias_v) b = __ldg(reinterpret_cast<const float4*>(bias_v + col));
      *reinterpret_cast<float4*>(t_agg + t * ldt + col) =
          make_float4(b.x * out_scale, b.y * out_scale, b.z * out_scale, b.w * out_scale);
    }
  }
}

// ---- register-resident q~ form (H <= 8): persistent, cp.async-pipelined, thread -> 4 columns of every head in BOTH phases ------
// ncu on the shared-memory form above (one token per CTA, synchronous copies): 27 % of the issue slots busy, warps stalled on
// the global loads of the next tile (long scoreboard 4.7 cycles per issue), on shared-memory operands (short scoreboard 5.5)
// and on the barriers around the copy -- a latency-bound kernel at 25 % of the HBM time it needs.  This form removes the exposed
// latencies instead of adding bandwidth:
//   * every thread keeps its 4 columns of all H transformed queries in registers and only ever touches ITS OWN 4 columns of a
//     centre row, in the score phase and in the weighted-sum phase alike, so rows are copied global -> shared with per-thread
//     cp.async (raw bytes; split-fp16 halves are joined when read) and need NO barrier: cp.async.wait_group is enough;
//   * CTAs are persistent (two per SM) and walk a stream of 8-row tiles that crosses token boundaries: while tile i is being
//     computed, tile i + 1 -- possibly the first tile of the NEXT token, together with that token's q~ -- is in flight;
//   * the 256-thread sum of the per-thread partial dot products is taken with a transposing butterfly: NV partial sums
//     (NV / H rows x H heads) are reduced over a warp with NV - 1 + log2(32 / NV) shuffles, after which lane l holds the warp
//     total of partial l % NV; the 8 warp totals meet in shared memory (one 16 B read per 4 H FMAs instead of two per 4).
constexpr int IB_ROWS = 8;           // rows per tile
constexpr int IB_STAGES = 2;

template <int N>
__device__ __forceinline__ void bfly_reduce(float* v, int lane) {     // N live values -> N / 2 ... -> 1: lane l then holds sum (l % N0)
  if constexpr (N >= 2) {
    constexpr int n = N / 2;
    const bool up = (lane & n) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, n);
    }
    bfly_reduce<n>(v, lane);
  }
}

// raw row bytes in shared memory: fp32 4 d, split fp16 4 d (hi | lo), bf16 2 d; `col` = first of the thread's 4 columns
template <typename InT>
__device__ __forceinline__ void ib_copy_row(char* dst_row, const InT* src_row, int col, int d) {
  if constexpr (std::is_same<InT, float>::value) {
    cp_async16(dst_row + col * 4, src_row + col);
  } else if constexpr (std::is_same<InT, __half>::value) {
    cp_async8(dst_row + col * 2, src_row + col);
    cp_async8(dst_row + (d + col) * 2, src_row + d + col);
  } else {
    cp_async8(dst_row + col * 2, src_row + col);
  }
}
template <typename InT>
__device__ __forceinline__ float4 ib_read_row(const char* row, int col, int d) {
  if constexpr (std::is_same<InT, float>::value) {
    return *reinterpret_cast<const float4*>(row + col * 4);
  } else if constexpr (std::is_same<InT, __half>::value) {
    return join4_f16(*reinterpret_cast<const uint2*>(row + col * 2), *reinterpret_cast<const uint2*>(row + (d + col) * 2));
  } else {
    const uint2 t = *reinterpret_cast<const uint2*>(row + col * 2);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
}

struct IbTile {            // one step of the tile stream: token, its edge range, first row of the tile
  int64_t tok;
  int64_t e0;
  int deg, r0;
};

template <typename InT, int H, int NV>
__global__ void __launch_bounds__(IA_THREADS, 2) inter_regq_kernel(const float* __restrict__ qt, int64_t q_hs,   // [H, T, d], head stride
                                                                const InT* __restrict__ hc, int64_t ldh,
                                                                const int32_t* __restrict__ indptr, int64_t t0, int64_t n_tokens,
                                                                int64_t d, __half* __restrict__ a_out, int64_t a_hs, int64_t lda,
                                                                const float* __restrict__ bias_v, float out_scale,
                                                                float* __restrict__ t_agg, int64_t ldt) {
  static_assert(NV % H == 0 && NV <= 32 && IB_ROWS % (NV / H) == 0, "NV / H rows per reduction group");
  constexpr int RG = NV / H;                               // rows per reduction group
  constexpr int NW = IA_THREADS / 32;
  constexpr int ROW_B = std::is_same<InT, __nv_bfloat16>::value ? 2 : 4;   // raw bytes per column
  extern __shared__ __align__(16) char ib_smem[];
  const int di = (int)d;
  const int row_bytes = di * ROW_B;
  char* rows = ib_smem;                                                    // [IB_STAGES][IB_ROWS][row_bytes], thread-private columns
  float* qbuf = reinterpret_cast<float*>(rows + IB_STAGES * IB_ROWS * row_bytes);   // [H][d] q~ of the token whose first tile is in flight
  float* wsum = qbuf + H * di;                             // [NW][H][IB_ROWS] warp totals of the tile's scores
  float* ss = wsum + NW * H * IB_ROWS;                     // [IB_ROWS][H] softmax numerators of the tile
  float* corr_s = ss + IB_ROWS * H;                        // [H] rescale of the running sums
  float* l_s = corr_s + H;                                 // [H] running denominators
  float* m_s = l_s + H;                                    // [H] running maxima

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int col = tid * 4;
  const bool active = col < di;

  // producer side of the tile stream (one tile ahead of the consumer); a token without centres is one empty tile
  auto token_tile = [&](int64_t tok) {
    IbTile it;
    it.tok = tok; it.r0 = 0; it.e0 = 0; it.deg = 0;
    if (tok < n_tokens) {
      it.e0 = __ldg(indptr + tok);
      it.deg = (int)(__ldg(indptr + tok + 1) - it.e0);
    }
    return it;
  };
  auto advance = [&](IbTile it) {
    it.r0 += IB_ROWS;
    if (it.r0 >= it.deg) it = token_tile(it.tok + gridDim.x);
    return it;
  };
  auto issue = [&](const IbTile& it, int stage) {           // cp.async of the tile (+ q~ when it opens a token); always one commit
    if (it.tok < n_tokens && active) {
      if (it.r0 == 0) {
#pragma unroll
        for (int h = 0; h < H; ++h) cp_async16(qbuf + h * di + col, qt + h * q_hs + (t0 + it.tok) * d + col);
      }
      const int nr = (it.deg - it.r0) < IB_ROWS ? (it.deg - it.r0) : IB_ROWS;
      char* dst = rows + stage * IB_ROWS * row_bytes;
      const InT* src = hc + (it.e0 + it.r0) * ldh;
#pragma unroll
      for (int r = 0; r < IB_ROWS; ++r)
        if (r < nr) ib_copy_row<InT>(dst + r * row_bytes, src + r * ldh, col, di);
    }
    cp_async_commit();
  };

  IbTile cur = token_tile(blockIdx.x);
  issue(cur, 0);
  float q[H][4], acc[H][4];
  int stage = 0;
  while (cur.tok < n_tokens) {
    const bool first = cur.r0 == 0;
    const int nr = (cur.deg - cur.r0) < IB_ROWS ? (cur.deg - cur.r0) : IB_ROWS;       // <= 0 for a token without centres
    const bool last = cur.r0 + IB_ROWS >= cur.deg;
    cp_async_wait_all();
    if (first) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) v = *reinterpret_cast<const float4*>(qbuf + h * di + col);
        q[h][0] = v.x; q[h][1] = v.y; q[h][2] = v.z; q[h][3] = v.w;
        acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;
      }
    }
    const IbTile nxt = advance(cur);
    issue(nxt, stage ^ 1);                                  // its buffer (and qbuf) were last read by this thread, before this point
    const char* my_rows = rows + stage * IB_ROWS * row_bytes;
    // ---- phase 1a: per-thread partial scores, reduced over the warp NV at a time
    for (int g0 = 0; g0 < nr; g0 += RG) {
      float v[NV];
#pragma unroll
      for (int rr = 0; rr < RG; ++rr) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && g0 + rr < nr) x = ib_read_row<InT>(my_rows + (g0 + rr) * row_bytes, col, di);
#pragma unroll
        for (int h = 0; h < H; ++h)
          v[rr * H + h] = fmaf(x.x, q[h][0], fmaf(x.y, q[h][1], fmaf(x.z, q[h][2], x.w * q[h][3])));
      }
      bfly_reduce<NV>(v, lane);
      float tot = v[0];
#pragma unroll
      for (int o = NV; o < 32; o <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane < NV) wsum[(warp * H + (lane % H)) * IB_ROWS + g0 + lane / H] = tot;
    }
    __syncthreads();
    // ---- phase 1b: block totals + online softmax bookkeeping (warp h, lane -> row)
    if (warp < H) {
      const int h = warp;
      float s = -INFINITY;
      if (lane < nr) {
        s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += wsum[(w * H + h) * IB_ROWS + lane];
      }
      const float m_old = first ? -INFINITY : m_s[h];
      const float l_old = first ? 0.f : l_s[h];
      const float mx = fmaxf(m_old, warp_max(s));
      const float p = lane < nr ? __expf(s - mx) : 0.f;
      const float tile_sum = warp_sum(p);
      __syncwarp();                                          // every lane has read m_s / l_s before lane 0 rewrites them
      if (lane < IB_ROWS) ss[lane * H + h] = p;
      if (lane == 0) {
        const float c = nr > 0 ? __expf(m_old - mx) : 0.f;   // first tile: exp(-inf) = 0
        corr_s[h] = c;
        l_s[h] = l_old * c + tile_sum;
        m_s[h] = mx;
      }
    }
    __syncthreads();
    // ---- phase 2: rescale and add the tile's weighted rows
    if (active) {
      if (!first) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float cr = corr_s[h];
          acc[h][0] *= cr; acc[h][1] *= cr; acc[h][2] *= cr; acc[h][3] *= cr;
        }
      }
#pragma unroll 2
      for (int r = 0; r < nr; ++r) {
        const float4 x = ib_read_row<InT>(my_rows + r * row_bytes, col, di);
        float pr[H];
#pragma unroll
        for (int h4 = 0; h4 < H / 4; ++h4) {
          const float4 p4 = *reinterpret_cast<const float4*>(ss + r * H + 4 * h4);
          pr[4 * h4] = p4.x; pr[4 * h4 + 1] = p4.y; pr[4 * h4 + 2] = p4.z; pr[4 * h4 + 3] = p4.w;
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
          acc[h][0] = fmaf(pr[h], x.x, acc[h][0]); acc[h][1] = fmaf(pr[h], x.y, acc[h][1]);
          acc[h][2] = fmaf(pr[h], x.z, acc[h][2]); acc[h][3] = fmaf(pr[h], x.w, acc[h][3]);
        }
      }
      if (last) {
        const int64_t t = t0 + cur.tok;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float inv = cur.deg > 0 ? 1.f / l_s[h] : 0.f;
          uint2 hi, lo;
          split4_f16(acc[h][0] * inv, acc[h][1] * inv, acc[h][2] * inv, acc[h][3] * inv, hi, lo);
          __half* o = a_out + h * a_hs + t * lda + col;
          *reinterpret_cast<uint2*>(o) = hi;
          *reinterpret_cast<uint2*>(o + d) = lo;
        }
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cur.deg > 0 && bias_v) b = __ldg(reinterpret_cast<const float4*>(bias_v + col));
        *reinterpret_cast<float4*>(t_agg + t * ldt + col) =
            make_float4(b.x * out_scale, b.y * out_scale, b.z * out_scale, b.w * out_scale);
      }
    }
    cur = nxt;
    stage ^= 1;
  }
  cp_async_wait_all();
}

// ---- tensor-core form (split-fp16 centre rows, H = 8, d = 128 KS): mma.sync m16n8k16, three fp16 passes --------------------------
// ncu on the register-q~ form: 170 M warp instructions per launch, 68 % of the issue slots busy -- the CUDA-core form is bound by
// instruction issue (64 FMAs per row per thread plus the butterfly and the hi + lo joins), 3x above the HBM time.  The centre rows
// already ARE fp16 pairs (hi | lo, the activation format of MATH_F16X3), and H = 8 is exactly the N of mma.m16n8k16, so both
// contractions go to the tensor cores with the same three-pass split the projections use (x_hi q_hi + x_hi q_lo + x_lo q_hi, fp32
// accumulation; the dropped x_lo q_lo term is 2^-22 relative):
//   scores   S[16 rows, 8 heads] = X[16, d] Q~^T      warp w contracts its d / 8 columns (A: ldmatrix of the row tile, B: the
//                                                     warp's slice of q~ as fp16 hi / lo fragments held in registers for the
//                                                     whole token, scaled by a power of two per (token, head) so that the lo
//                                                     halves stay normal); the 8 warp partials meet in shared memory
//   sums     A^T[d, 8 heads] += X^T[d, 16] P[16, 8]   warp w owns its d / 8 columns of all heads (A: ldmatrix.trans of the same
//                                                     tile, B: the tile's softmax numerators x 1024 as fp16 hi / lo)
// Same persistent tile stream as above (cp.async, one 16-row tile in flight under the math of the previous one, crossing
// token boundaries), one CTA per SM.  ~100 warp instructions per tile and warp instead of ~2000.
constexpr int IM_ROWS = 16;

constexpr int IM_STAGES = 3;         // two 16-row tiles (128 KB at d = 1024) in flight per SM under the math of a third
constexpr int IM_TAB = 512;          // tokens per CTA and launch whose edge ranges are staged in shared memory

// Two block barriers per tile: (1) tile landed / previous tile consumed, (2) the 8 per-warp partial score tiles are in shared
// memory.  Everything after that is per warp: every warp sums the partials, runs the online softmax of all 8 heads redundantly
// with lane (g, tq) owning head g and rows {2 tq, 2 tq + 1, 2 tq + 8, 2 tq + 9} -- exactly its B fragment of P -- and keeps the
// running maxima / denominators in registers; q~ is scaled per (warp, head); the normalised sums leave the accumulator fragments
// as 2-byte stores (8 consecutive lanes cover 16 contiguous bytes of a hi or lo row).
// BF16: centre rows are bf16 (MATH_BF16 activations) -- one A operand, q~ and P split into bf16 hi / lo (two passes, no scaling:
// bf16 has the fp32 exponent range); otherwise split fp16 rows, three passes.
// NW warps per CTA (8 or 16): with 16, a warp owns d / 16 columns -- half the tensor work and accumulator registers per warp, twice
// the warps per scheduler to hide the per-tile dependency chain (ldmatrix -> mma -> barrier -> softmax -> mma).
template <int KS, bool BF16, int NW>
__global__ void __launch_bounds__(NW * 32, 1) inter_mma_kernel(const float* __restrict__ qt, int64_t q_hs,   // [8, T, d], head stride
                                                               const void* __restrict__ hc_, int64_t ldh,
                                                               const int32_t* __restrict__ indptr, int64_t t0, int64_t n_tokens,
                                                               __half* __restrict__ a_out, int64_t a_hs, int64_t lda,
                                                               const float* __restrict__ bias_v, float out_scale,
                                                               float* __restrict__ t_agg, int64_t ldt, int bulk) {
  constexpr int H = 8, D = 128 * KS, NT = NW * 32;
  constexpr int KW = KS * 8 / NW;                          // 16-column k-steps of a warp
  static_assert(KW * NW == KS * 8, "d / 16 must divide by the number of warps");
  constexpr int ROWB = BF16 ? 2 * D : 4 * D;               // bytes of one centre row (bf16, or fp16 hi | lo)
  constexpr int RS = ROWB + 16;                            // row stride in shared memory (+ 16: conflict-free ldmatrix)
  const char* hc = reinterpret_cast<const char*>(hc_);
  extern __shared__ __align__(128) char im_smem[];
  char* rows = im_smem;                                                  // [IM_STAGES][16][RS]
  float* wsum = reinterpret_cast<float*>(rows + IM_STAGES * IM_ROWS * RS);   // [NW][16][8] per-warp partial scores
  float* isc_s = wsum + NW * IM_ROWS * H;                                // [NW][8] 1 / (power-of-two scale of the warp's q~ slice)
  int2* tab = reinterpret_cast<int2*>(isc_s + NW * H);                   // [IM_TAB] (first edge, degree) of this CTA's tokens
  constexpr int OSW = KW * 16 + 4;                                       // staging row stride in floats
  float* ostage = reinterpret_cast<float*>(tab + IM_TAB);                // [NW][4][OSW] per-warp output staging

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const bool active = tid * 4 < D;                         // owner of 4 output columns (bias store)
  const int chunk = tid & 255, rpar = tid >> 8;            // 16 B chunk of a row; rows rpar, rpar + NT / 256, ...
  const bool copier = chunk * 16 < ROWB;
  const int wcol0 = warp * KW * 16;                        // first of the warp's columns
  const int mi = lane >> 3, r8 = lane & 7;
  const int p1_off = (r8 + (mi & 1) * 8) * RS + (mi >> 1) * 16 + wcol0 * 2;     // A = X      (rows x columns)
  const int p2_off = (r8 + (mi >> 1) * 8) * RS + (mi & 1) * 16 + wcol0 * 2;     // A = X^T    (columns x rows), ldmatrix.trans

  // edge ranges of this CTA's tokens (token j of the CTA = blockIdx.x + j gridDim.x): one strided read up front instead of a
  // dependent global load at every token switch (ncu: 12 % of the stall samples)
  const int n_mine = n_tokens > blockIdx.x ? (int)((n_tokens - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  for (int j = tid; j < n_mine; j += NT) {
    const int64_t tok = blockIdx.x + (int64_t)j * gridDim.x;
    const int e0 = __ldg(indptr + tok);
    tab[j] = make_int2(e0, __ldg(indptr + tok + 1) - e0);
  }
  // `bulk`: the rows of a tile arrive by one cp.async.bulk each (TMA engine, mbarrier complete_tx) instead of 256 x 16 B cp.async per
  // row -- no LSU instruction per 16 B (ncu on the cp.async form: MIO-throttle / short-scoreboard stalls on the load issue)
  __shared__ __align__(8) uint64_t full_bar[IM_STAGES];
  if (bulk && tid == 0) {
    for (int s_ = 0; s_ < IM_STAGES; ++s_) ms_mbar_init(&full_bar[s_], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  struct Tile { int j, e0, deg, r0; };                     // token j of this CTA, first row of the tile
  auto token_tile = [&](int j) {
    Tile it;
    it.j = j; it.r0 = 0; it.e0 = 0; it.deg = 0;
    if (j < n_mine) {
      const int2 e = tab[j];
      it.e0 = e.x; it.deg = e.y;
    }
    return it;
  };
  auto advance = [&](Tile it) {
    it.r0 += IM_ROWS;
    if (it.r0 >= it.deg) it = token_tile(it.j + 1);
    return it;
  };
  auto issue = [&](const Tile& it, int st) {               // cp.async of one tile; always exactly one commit
    if (bulk) {
      if (it.j < n_mine) {
        const int nr = (it.deg - it.r0) < IM_ROWS ? (it.deg - it.r0) : IM_ROWS;
        char* dst = rows + st * IM_ROWS * RS;
        if (warp == 0) {
          if (lane == 0) ms_mbar_expect_tx(&full_bar[st], (uint32_t)(nr > 0 ? nr : 0) * ROWB);
          __syncwarp();
          if (lane < nr) ms_bulk_g2s(dst + lane * RS, hc + (int64_t)(it.e0 + it.r0 + lane) * ldh * 2, ROWB, &full_bar[st]);
        } else {                                           // rows past the token's last centre: finite (zero)
          for (int r = (nr > 0 ? nr : 0) + (warp - 1); r < IM_ROWS; r += NW - 1)
            for (int c = lane * 16; c < ROWB; c += 512) *reinterpret_cast<uint4*>(dst + r * RS + c) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      return;
    }
    if (it.j < n_mine && copier) {
      const int nr = (it.deg - it.r0) < IM_ROWS ? (it.deg - it.r0) : IM_ROWS;
      char* dst = rows + st * IM_ROWS * RS + chunk * 16;
      const char* src = hc + (int64_t)(it.e0 + it.r0) * ldh * 2 + chunk * 16;
#pragma unroll
      for (int r = rpar; r < IM_ROWS; r += NT / 256) {
        if (r < nr) cp_async16(dst + r * RS, src + (int64_t)r * ldh * 2);
        else *reinterpret_cast<uint4*>(dst + r * RS) = make_uint4(0u, 0u, 0u, 0u);       // rows past the token's last centre: finite
      }
    }
    cp_async_commit();
  };
  // this thread's B-fragment elements of q~ (head g, columns wcol0 + 16 ks + 2 tq + {0, 1, 8, 9}), raw fp32, loaded one token ahead
  float2 qraw[KW][2];
  auto load_q = [&](int j) {
    if (j < n_mine) {
      const float* qrow = qt + g * q_hs + (t0 + blockIdx.x + (int64_t)j * gridDim.x) * D + wcol0 + 2 * tq;
#pragma unroll
      for (int ks = 0; ks < KW; ++ks) {
        qraw[ks][0] = __ldg(reinterpret_cast<const float2*>(qrow + ks * 16));
        qraw[ks][1] = __ldg(reinterpret_cast<const float2*>(qrow + ks * 16 + 8));
      }
    }
  };

  static_assert(IM_STAGES == 3, "cur, n1, n2: the tile being computed and the two in flight");
  Tile cur = token_tile(0);
  load_q(0);
  Tile n1 = advance(cur);
  issue(cur, 0);
  issue(n1, 1);
  Tile n2 = advance(n1);                                   // next tile to issue
  uint32_t qh[KW][2], ql[KW][2];
  float acc[KW][4];
  float m_run = -INFINITY, l_run = 0.f;                    // head g (replicated over tq and over the warps)
  float isc[NW > 8 ? 1 : NW];                              // 1 / scale of head g's q~ slice in every warp (NW > 8: read from shared memory)
  int stage = 0;
  uint32_t tile_no = 0;
  while (cur.j < n_mine) {
    const bool first = cur.r0 == 0;
    const int nr = (cur.deg - cur.r0) < IM_ROWS ? (cur.deg - cur.r0) : IM_ROWS;        // 0 for a token without centres
    const bool last = cur.r0 + IM_ROWS >= cur.deg;
    if (bulk) ms_mbar_wait(&full_bar[stage], (tile_no / IM_STAGES) & 1);
    else asm volatile("cp.async.wait_group %0;" ::"n"(IM_STAGES - 2) : "memory");
    ++tile_no;
    __syncthreads();                                       // tile visible; everyone is past the previous tile
    issue(n2, stage == 0 ? IM_STAGES - 1 : stage - 1);     // into the buffer of the previous tile
    const Tile n3 = advance(n2);
    if (first) {
      // q~ slice of this warp -> fp16 hi / lo B fragments, scaled per (warp, head) by a power of two (amax -> [4096, 8192])
      float am = 0.f;
#pragma unroll
      for (int ks = 0; ks < KW; ++ks)
        am = fmaxf(fmaxf(am, fmaxf(fabsf(qraw[ks][0].x), fabsf(qraw[ks][0].y))), fmaxf(fabsf(qraw[ks][1].x), fabsf(qraw[ks][1].y)));
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
      float scale = 1.f;
      if (!BF16 && am > 0.f && am < INFINITY) scale = exp2f(fminf(fmaxf(floorf(log2f(8192.f / am)), -100.f), 100.f));
      if (tq == 0) isc_s[warp * H + g] = 1.f / scale;
#pragma unroll
      for (int ks = 0; ks < KW; ++ks) {
        if constexpr (BF16) {
          split2_bf16(qraw[ks][0].x, qraw[ks][0].y, qh[ks][0], ql[ks][0]);
          split2_bf16(qraw[ks][1].x, qraw[ks][1].y, qh[ks][1], ql[ks][1]);
        } else {
          split2_f16(qraw[ks][0].x * scale, qraw[ks][0].y * scale, qh[ks][0], ql[ks][0]);
          split2_f16(qraw[ks][1].x * scale, qraw[ks][1].y * scale, qh[ks][1], ql[ks][1]);
        }
        acc[ks][0] = acc[ks][1] = acc[ks][2] = acc[ks][3] = 0.f;
      }
      m_run = -INFINITY;
      l_run = 0.f;
      load_q(cur.j + 1);                                   // in flight for the whole token
    }
    const char* tile = rows + stage * IM_ROWS * RS;
    // ---- phase 1a: this warp's partial scores of the tile (tensor cores; one accumulator per pass: three independent chains)
    {
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KW; ++ks) {
        uint32_t ah[4];
        ldsm_x4(ah, tile + p1_off + ks * 32);
        if constexpr (BF16) {
          mma16816_bf16(c0, ah, qh[ks][0], qh[ks][1]);
          mma16816_bf16(c1, ah, ql[ks][0], ql[ks][1]);
        } else {
          uint32_t al[4];
          ldsm_x4(al, tile + p1_off + ks * 32 + D * 2);
          mma16816(c0, ah, qh[ks][0], qh[ks][1]);
          mma16816(c1, ah, ql[ks][0], ql[ks][1]);
          mma16816(c2, al, qh[ks][0], qh[ks][1]);
        }
      }
      *reinterpret_cast<float2*>(wsum + (warp * IM_ROWS + g) * H + 2 * tq) = make_float2(c0[0] + (c1[0] + c2[0]), c0[1] + (c1[1] + c2[1]));
      *reinterpret_cast<float2*>(wsum + (warp * IM_ROWS + g + 8) * H + 2 * tq) =
          make_float2(c0[2] + (c1[2] + c2[2]), c0[3] + (c1[3] + c2[3]));
    }
    __syncthreads();
    // ---- phase 1b (every warp): scores of head g at rows 2 tq + {0, 1, 8, 9}, online softmax, P fragments
    if constexpr (NW <= 8) {
      if (first) {
#pragma unroll
        for (int w = 0; w < NW; ++w) isc[w] = isc_s[w * H + g];
      }
    }
    float sc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = 2 * tq + (j & 1) + (j >> 1) * 8;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s = fmaf(wsum[(w * IM_ROWS + row) * H + g], NW <= 8 ? isc[NW <= 8 ? w : 0] : isc_s[w * H + g], s);
      sc[j] = row < nr ? s : -INFINITY;
    }
    float mx = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = fmaxf(mx, m_run);
    float pj[4], tile_sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pj[j] = sc[j] > -INFINITY ? __expf(sc[j] - mx) : 0.f;
      tile_sum += pj[j];
    }
    tile_sum += __shfl_xor_sync(0xffffffffu, tile_sum, 1);
    tile_sum += __shfl_xor_sync(0xffffffffu, tile_sum, 2);
    const float corr = nr > 0 && m_run > -INFINITY ? __expf(m_run - mx) : 0.f;
    l_run = l_run * corr + tile_sum;
    m_run = mx;
    // ---- phase 2: rescale, add the tile's weighted rows (tensor cores); accumulator (column g / g + 8 of tile mt, heads 2 tq, 2 tq + 1)
    {
      const float cr0 = __shfl_sync(0xffffffffu, corr, 8 * tq), cr1 = __shfl_sync(0xffffffffu, corr, 8 * tq + 4);
      if (!first) {
#pragma unroll
        for (int mt = 0; mt < KW; ++mt) {
          acc[mt][0] *= cr0; acc[mt][1] *= cr1; acc[mt][2] *= cr0; acc[mt][3] *= cr1;
        }
      }
    }
    if (nr > 0) {
      uint32_t bh0, bl0, bh1, bl1;
      if constexpr (BF16) {
        split2_bf16(pj[0] * 1024.f, pj[1] * 1024.f, bh0, bl0);
        split2_bf16(pj[2] * 1024.f, pj[3] * 1024.f, bh1, bl1);
      } else {
        split2_f16(pj[0] * 1024.f, pj[1] * 1024.f, bh0, bl0);
        split2_f16(pj[2] * 1024.f, pj[3] * 1024.f, bh1, bl1);
      }
#pragma unroll
      for (int mt = 0; mt < KW; ++mt) {
        uint32_t ah[4];
        ldsm_x4_t(ah, tile + p2_off + mt * 32);
        if constexpr (BF16) {
          mma16816_bf16(acc[mt], ah, bl0, bl1);
          mma16816_bf16(acc[mt], ah, bh0, bh1);
        } else {
          uint32_t al[4];
          ldsm_x4_t(al, tile + p2_off + mt * 32 + D * 2);
          mma16816(acc[mt], al, bh0, bh1);
          mma16816(acc[mt], ah, bl0, bl1);
          mma16816(acc[mt], ah, bh0, bh1);
        }
      }
    }
    if (last) {
      const float l0 = __shfl_sync(0xffffffffu, l_run, 8 * tq), l1 = __shfl_sync(0xffffffffu, l_run, 8 * tq + 4);
      const float i0 = cur.deg > 0 ? 1.f / (1024.f * l0) : 0.f;
      const float i1 = cur.deg > 0 ? 1.f / (1024.f * l1) : 0.f;
      const int64_t t = t0 + blockIdx.x + (int64_t)cur.j * gridDim.x;
      // the warp's [8 heads x 16 KS columns] block leaves through a private staging area, four heads at a time (lanes with
      // tq < 2 hold heads 0-3), as coalesced split-fp16 row segments: no block barrier
      float* stg = ostage + warp * (4 * OSW);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        __syncwarp();
        if ((tq >> 1) == half) {
#pragma unroll
          for (int mt = 0; mt < KW; ++mt) {
            float* o = stg + (2 * (tq & 1)) * OSW + mt * 16 + g;
            o[0] = acc[mt][0] * i0; o[OSW] = acc[mt][1] * i1; o[8] = acc[mt][2] * i0; o[OSW + 8] = acc[mt][3] * i1;
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < (4 * KW * 16 + 127) / 128; ++i) {            // 4 heads x 16 KW columns, 4 per lane and step
          const int idx = i * 128 + lane * 4;
          if (idx >= 4 * KW * 16) break;
          const int hh = idx / (KW * 16), cc = idx % (KW * 16);
          const float4 v = *reinterpret_cast<const float4*>(stg + hh * OSW + cc);
          uint2 hi, lo;
          split4_f16(v.x, v.y, v.z, v.w, hi, lo);
          __half* o = a_out + (4 * half + hh) * a_hs + t * lda + wcol0 + cc;
          *reinterpret_cast<uint2*>(o) = hi;
          *reinterpret_cast<uint2*>(o + D) = lo;
        }
      }
      if (active) {
        const int col = tid * 4;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cur.deg > 0 && bias_v) b = __ldg(reinterpret_cast<const float4*>(bias_v + col));
        *reinterpret_cast<float4*>(t_agg + t * ldt + col) =
            make_float4(b.x * out_scale, b.y * out_scale, b.z * out_scale, b.w * out_scale);
      }
    }
    cur = n1; n1 = n2; n2 = n3;
    stage = stage + 1 == IM_STAGES ? 0 : stage + 1;
  }
  cp_async_wait_all();
}

template <int KS, bool BF16, int NW>
static int32_t launch_inter_mma(const float* qt, int64_t q_hs, const void* hc, int64_t ldh, const int32_t* indptr, int64_t t0,
                                int64_t n_tokens, void* a_out, int64_t a_hs, int64_t lda, const float* bias_v, float out_scale,
                                float* t_agg, int64_t ldt, int n_sm, cudaStream_t st) {
  constexpr int D = 128 * KS, KW = KS * 8 / NW;
  const size_t smem = (size_t)IM_STAGES * IM_ROWS * ((BF16 ? 2 : 4) * D + 16) + ((size_t)NW * IM_ROWS * 8 + NW * 8) * sizeof(float) +
                      (size_t)IM_TAB * sizeof(int2) + (size_t)NW * 4 * (KW * 16 + 4) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    GNNLM_CUDA(cudaFuncSetAttribute(inter_mma_kernel<KS, BF16, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  // persistent: one CTA per SM; a launch covers at most IM_TAB tokens per CTA (their edge ranges are staged in shared memory)
  const int64_t per_launch = (int64_t)n_sm * IM_TAB;
  for (int64_t b0 = 0; b0 < n_tokens; b0 += per_launch) {
    const int64_t n = n_tokens - b0 < per_launch ? n_tokens - b0 : per_launch;
    const int64_t grid = n < n_sm ? n : n_sm;
    // GNNLM_INTER_BULK=0: cp.async tile loads (the round-1 form) instead of one bulk copy per row (A/B timing switch)
    static const int bulk = [] { const char* e = getenv("GNNLM_INTER_BULK"); return e ? atoi(e) : 1; }();
    inter_mma_kernel<KS, BF16, NW><<<(unsigned)grid, NW * 32, smem, st>>>(qt, q_hs, hc, ldh, indptr + b0, t0 + b0, n, (__half*)a_out,
                                                                         a_hs, lda, bias_v, out_scale, t_agg, ldt, bulk);
    GNNLM_LAUNCH_CHECK("gnnlm_hgt_inter_fused");
  }
  return 0;
}

template <typename InT, int H>
static int32_t launch_inter(const float* qt, int64_t q_hs, const void* hc, int64_t ldh, const int32_t* indptr, int64_t t0,
                            int64_t n_tokens, int64_t d, void* a_out, int64_t a_hs, int64_t lda, const float* bias_v,
                            float out_scale, float* t_agg, int64_t ldt, cudaStream_t st) {
  // GNNLM_INTER_KERNEL=smemq | regq selects the shared-memory q~ / register q~ CUDA-core forms where the tensor-core form
  // would run (A/B timing switch)
  static const int force = [] {
    const char* e = getenv("GNNLM_INTER_KERNEL");
    return !e ? 0 : strcmp(e, "smemq") == 0 ? 1 : strcmp(e, "regq") == 0 ? 2 : 0;
  }();
  const bool force_smemq = force == 1;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    GNNLM_CUDA(cudaGetDevice(&dev));
    GNNLM_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  if constexpr (!std::is_same<InT, float>::value && H == 8) {
    if (force == 0 && ldh % 8 == 0 && d % 128 == 0) {
      // GNNLM_INTER_WARPS=16: the 16-warp form (A/B timing switch).  Measured on the Wiki103 shape (profiles/r2_inter_probe.log):
      // 8 warps 156 us (3.93 TB/s), 16 warps 197 us -- the softmax every warp repeats and the wider barriers cost more than
      // the extra warps hide, so 8 stays the default
      static const int nw = [] { const char* e = getenv("GNNLM_INTER_WARPS"); return e && atoi(e) == 16 ? 16 : 8; }();
#define GNNLM_IM(KS)                                                                                                              \
  return nw == 8 ? launch_inter_mma<KS, std::is_same<InT, __nv_bfloat16>::value, 8>(qt, q_hs, hc, ldh, indptr, t0, n_tokens, a_out, a_hs, \
                                                                                  lda, bias_v, out_scale, t_agg, ldt, n_sm, st)   \
                 : launch_inter_mma<KS, std::is_same<InT, __nv_bfloat16>::value, 16>(qt, q_hs, hc, ldh, indptr, t0, n_tokens, a_out,      \
                                                                                   a_hs, lda, bias_v, out_scale, t_agg, ldt, n_sm, st)
      switch (d / 128) {
        case 4: GNNLM_IM(4);
        case 6: GNNLM_IM(6);
        case 8: GNNLM_IM(8);
        default: break;
      }
#undef GNNLM_IM
    }
  }
  if constexpr (H <= 8) {
    if (!force_smemq) {
      constexpr int NV = 16;                                 // partial sums per butterfly (16: 2 rows x 8 heads / 4 rows x 4 heads)
      const size_t row_b = (size_t)d * (std::is_same<InT, __nv_bfloat16>::value ? 2 : 4);
      const size_t smem = (size_t)IB_STAGES * IB_ROWS * row_b +
                          ((size_t)H * d + (size_t)(IA_THREADS / 32) * H * IB_ROWS + IB_ROWS * H + 3 * H) * sizeof(float);
      static size_t configured = 0;
      if (smem > configured) {
        GNNLM_CUDA(cudaFuncSetAttribute(inter_regq_kernel<InT, H, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
      }
      const int64_t grid = n_tokens < 2 * (int64_t)n_sm ? n_tokens : 2 * (int64_t)n_sm;      // persistent: two CTAs per SM
      inter_regq_kernel<InT, H, NV><<<(unsigned)grid, IA_THREADS, smem, st>>>(qt, q_hs, (const InT*)hc, ldh, indptr, t0, n_tokens, d,
                                                                             (__half*)a_out, a_hs, lda, bias_v, out_scale, t_agg,
                                                                             ldt);
      GNNLM_LAUNCH_CHECK("gnnlm_hgt_inter_fused");
      return 0;
    }
  }
  const size_t smem = ((size_t)IA_ROWS * (d + 4) + (size_t)H * d + IA_ROWS * H + 3 * H) * sizeof(float);
  GNNLM_CHECK_ARG(smem <= 220 * 1024, GNNLM_E_UNSUPPORTED, "gnnlm_hgt_inter_fused: H * d too large for shared memory (%zu B)", smem);
  static size_t configured = 0;
  if (smem > configured) {
    GNNLM_CUDA(cudaFuncSetAttribute(inter_fused_kernel<InT, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  inter_fused_kernel<InT, H><<<(unsigned)n_tokens, IA_THREADS, smem, st>>>(qt, q_hs, (const InT*)hc, ldh, indptr, t0, d,
                                                                          (__half*)a_out, a_hs, lda, bias_v, out_scale, t_agg,
                                                                          ldt);
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_inter_fused");
  return 0;
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_hgt_inter_fused(const float* q_tilde, int64_t q_head_stride, const void* hc, int32_t hc_dtype,
                                         int64_t ldh, const int32_t* inter_indptr, int64_t t0, int64_t n_tokens, int32_t H,
                                         int64_t d, void* a_out, int64_t a_head_stride, int64_t lda, const float* bias_v,
                                         float out_scale, float* t_agg, int64_t ldt, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(q_tilde && hc && inter_indptr && a_out && t_agg, GNNLM_E_ARG, "gnnlm_hgt_inter_fused: null pointer");
  GNNLM_CHECK_ARG(hc_dtype == GNNLM_F32 || hc_dtype == GNNLM_BF16 || hc_dtype == GNNLM_F16X2, GNNLM_E_UNSUPPORTED,
                  "gnnlm_hgt_inter_fused: centre features must be F32, BF16 or F16X2");
  GNNLM_CHECK_ARG(d > 0 && d % 8 == 0 && d <= IA_THREADS * 4 && ldh % 4 == 0 && lda % 4 == 0 && lda >= 2 * d && ldt % 4 == 0 &&
                      ldt >= d && (uintptr_t)q_tilde % 16 == 0 && (uintptr_t)hc % 16 == 0 && (uintptr_t)a_out % 16 == 0 &&
                      (uintptr_t)t_agg % 16 == 0 && q_head_stride % 4 == 0 && a_head_stride % 4 == 0,
                  GNNLM_E_SHAPE, "gnnlm_hgt_inter_fused: d must be a multiple of 8 and <= %d, strides multiples of 4, pointers 16 B aligned",
                  IA_THREADS * 4);
  GNNLM_CHECK_ARG(n_tokens >= 0 && t0 >= 0, GNNLM_E_SHAPE, "gnnlm_hgt_inter_fused: bad token range");
  if (n_tokens == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define GNNLM_IA(T, HH)                                                                                                       \
  return launch_inter<T, HH>(q_tilde, q_head_stride, hc, ldh, inter_indptr, t0, n_tokens, d, a_out, a_head_stride, lda, bias_v, \
                             out_scale, t_agg, ldt, st)
#define GNNLM_IA_H(T)                      \
  switch (H) {                             \
    case 4: GNNLM_IA(T, 4);                \
    case 8: GNNLM_IA(T, 8);                \
    case 12: GNNLM_IA(T, 12);              \
    case 16: GNNLM_IA(T, 16);              \
    default: break;                        \
  }
  if (hc_dtype == GNNLM_F32) { GNNLM_IA_H(float) }
  else if (hc_dtype == GNNLM_F16X2) { GNNLM_IA_H(__half) }
  else { GNNLM_IA_H(__nv_bfloat16) }
#undef GNNLM_IA_H
#undef GNNLM_IA
  gnnlm::set_error("gnnlm_hgt_inter_fused: H must be 4, 8, 12 or 16 (H=%d)", H);
  return GNNLM_E_UNSUPPORTED;
}
