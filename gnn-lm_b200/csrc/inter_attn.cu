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
// i.e. two per-head GEMMs over the T target rows (gemm_tcgen05.cu, batched) around THIS kernel, which per token streams
// its <= k centre rows once, scores them against the H transformed queries, and writes the H alpha-weighted row sums:
// d reads + H*d writes per token instead of 2 d^2 MACs per centre.  Same result up to fp32 re-association.
//
// One CTA (256 threads) per token, two CTAs per SM; centre rows in tiles of 16 through shared memory (fp32, row stride
// d + 4 floats: conflict-free for both access patterns), online softmax across tiles.  Phase 1: warp -> head,
// lane -> (half of the columns, row).  Phase 2: thread -> 4 columns of all H sums.
#include <type_traits>

#include "common.cuh"

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
#pragma unroll
        for (int h = 0; h < HH; ++h) {
          const float p = ss[r * H + h0 + h];
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

template <typename InT, int H>
static int32_t launch_inter(const float* qt, int64_t q_hs, const void* hc, int64_t ldh, const int32_t* indptr, int64_t t0,
                            int64_t n_tokens, int64_t d, void* a_out, int64_t a_hs, int64_t lda, const float* bias_v,
                            float out_scale, float* t_agg, int64_t ldt, cudaStream_t st) {
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
