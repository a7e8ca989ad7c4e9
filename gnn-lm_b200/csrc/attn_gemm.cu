// Helpers that let the ('tgt','intra','tgt') causal attention of HGTLayer.forward (reference:
// fairseq/models/hgt.py:354-358 over the edges of token_block_dataset.py:586-594) run on the tensor cores at
// fp32 parity: per head, S = Q K'^T and O = softmax_causal(S) V' are two 3xFP16 GEMMs (gemm_tcgen05.cu,
// gemm_f16s_kernel) and everything in between is the three HBM-streaming kernels below.
//
//   heads_split      Q / K' [L, H*d_k] fp32 (row stride ld) -> per-head split-fp16 operands:
//                    A-style [H, L, 2*d_k] (hi | lo in one row) or W-style hi [H, L, d_k], lo [H, L, d_k]
//   heads_transpose  V' [L, H*d_k] -> W-style transposed operands hi / lo [H, d_k, L]  (P V' = P (V'^T)^T)
//   causal_softmax   S [H, L, L] fp32 -> P split-fp16 [H, L, 2L]: row i is softmax over j in [max(0, i-ctx+1), i],
//                    zero elsewhere (the masked products then contribute exact zeros to the GEMM)
//
// Both GEMMs run with causal tile scheduling (gnnlm_linear_batched_f16x3 `causal`): S tiles entirely above the diagonal are
// never computed (nor read here), and P V' contracts row pair r only over k < (r+1)*256, so P is written up to there only.
//
// The lower-triangular tiles of the L x L score matrix are materialised per block (H*L*L*4 B = 302 MB allocated at
// L = 3072, H = 8, ~54 % of it touched), still ~3x faster than the fp32 CUDA-core flash kernel.
#include <stdlib.h>

#include "common.cuh"

namespace gnnlm {

__global__ void __launch_bounds__(256) heads_split_kernel(const float* __restrict__ src, int64_t ld, int64_t L, int H, int dk,
                                                          int a_style, __half* __restrict__ hi, __half* __restrict__ lo) {
  // one thread per 4 consecutive features of one (token, head)
  const int64_t per_row = (int64_t)H * dk / 4;
  const int64_t n = L * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / per_row;
    const int c = (int)(i % per_row) * 4;
    const int h = c / dk, j = c % dk;
    const float4 x = __ldg(reinterpret_cast<const float4*>(src + t * ld + c));
    uint2 ph, pl;
    split4_f16(x.x, x.y, x.z, x.w, ph, pl);
    if (a_style) {
      __half* row = hi + ((int64_t)h * L + t) * 2 * dk;
      *reinterpret_cast<uint2*>(row + j) = ph;
      *reinterpret_cast<uint2*>(row + dk + j) = pl;
    } else {
      const int64_t o = ((int64_t)h * L + t) * dk + j;
      *reinterpret_cast<uint2*>(hi + o) = ph;
      *reinterpret_cast<uint2*>(lo + o) = pl;
    }
  }
}

// 32 x 32 smem tile transpose per (head, token tile, feature tile)
__global__ void __launch_bounds__(256) heads_transpose_kernel(const float* __restrict__ src, int64_t ld, int64_t L, int H, int dk,
                                                              __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int h = blockIdx.z;
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const int j0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;             // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t t = t0 + r;
    tile[r][tx] = (t < L && j0 + tx < dk) ? __ldg(src + t * ld + (int64_t)h * dk + j0 + tx) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {                                   // r = feature inside the tile, tx = token
    const int64_t t = t0 + tx;
    const int j = j0 + r;
    if (t < L && j < dk) {
      const float x = fminf(fmaxf(tile[tx][r], -65504.f), 65504.f);
      const __half a = __float2half_rn(x);
      const int64_t o = ((int64_t)h * dk + j) * L + t;
      hi[o] = a;
      lo[o] = __float2half_rn(x - __half2float(a));
    }
  }
}

// one warp per (head, query row), four columns per lane per step (L % 4 == 0): pass 1 keeps an online (max, sum) per lane,
// pass 2 writes hi | lo quads.  Columns j > i are never read (their tiles may be unwritten).
__global__ void __launch_bounds__(256) causal_softmax_kernel(const float* __restrict__ S, int64_t L, int64_t ctx, int H,
                                                             int64_t k_tile, __half* __restrict__ P) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)H * L;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const int64_t i = r % L;
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
      if (mx > m) {                       // rescale the running sum (m == -inf: l is still 0)
        l *= __expf(m - mx);
        m = mx;
      }
      if (m > -INFINITY) {
#pragma unroll
        for (int e = 0; e < 4; ++e) l += __expf(x[e] - m);
      }
    }
    const float M = warp_max(m);          // the diagonal is always valid, so M is finite
    l = warp_sum(m > -INFINITY ? l * __expf(m - M) : 0.f);
    const float inv = 1.f / l;
    // the consumer (P V' with causal == 2) contracts row i over k < (i / k_tile + 1) * k_tile only
    const int64_t j_end = k_tile > 0 ? min(L, (i / k_tile + 1) * k_tile) : L;
    for (int64_t j = lane * 4; j < j_end; j += 128) {
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      if (j <= i && j + 3 >= lo_j) {
        const float4 x4 = *reinterpret_cast<const float4*>(s + j);
        const float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j + e >= lo_j && j + e <= i) w[e] = __expf(x[e] - M) * inv;
      }
      uint2 hi, lo;
      split4_f16(w[0], w[1], w[2], w[3], hi, lo);
      *reinterpret_cast<uint2*>(p + j) = hi;
      *reinterpret_cast<uint2*>(p + L + j) = lo;
    }
  }
}

// Register-resident form for L <= 128 * NT: the row is read ONCE (NT float4 per lane, all loads in flight together), max / sum / the
// normalised numerators come from registers -- one exponential per logit instead of two plus the online rescales, and no second
// pass over S.  Same arithmetic as above up to the order of the (max, sum) reduction.
template <int NT>
__global__ void __launch_bounds__(256) causal_softmax_reg_kernel(const float* __restrict__ S, int64_t L, int64_t ctx, int H,
                                                                 int64_t k_tile, __half* __restrict__ P) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)H * L;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const int i = (int)(r % L);
    const float* s = S + r * L;
    __half* p = P + r * 2 * L;
    const int lo_j = (ctx > 0 && i + 1 > ctx) ? (int)(i + 1 - ctx) : 0;
    float4 x[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int j = t * 128 + lane * 4;
      x[t] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (j <= i && j + 3 >= lo_j) {                      // columns j > i are never read (their tiles may be unwritten)
        x[t] = __ldg(reinterpret_cast<const float4*>(s + j));
        if (j < lo_j) x[t].x = -INFINITY;
        if (j + 1 < lo_j || j + 1 > i) x[t].y = -INFINITY;
        if (j + 2 < lo_j || j + 2 > i) x[t].z = -INFINITY;
        if (j + 3 < lo_j || j + 3 > i) x[t].w = -INFINITY;
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < NT; ++t) m = fmaxf(m, fmaxf(fmaxf(x[t].x, x[t].y), fmaxf(x[t].z, x[t].w)));
    const float M = warp_max(m);                          // the diagonal is always valid, so M is finite
    float l = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      x[t].x = __expf(x[t].x - M); x[t].y = __expf(x[t].y - M); x[t].z = __expf(x[t].z - M); x[t].w = __expf(x[t].w - M);
      l += (x[t].x + x[t].y) + (x[t].z + x[t].w);         // exp(-inf) = 0 for masked columns
    }
    const float inv = 1.f / warp_sum(l);
    // the consumer (P V' with causal == 2) contracts row i over k < (i / k_tile + 1) * k_tile only
    const int j_end = (int)(k_tile > 0 ? min(L, (i / k_tile + 1) * k_tile) : L);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int j = t * 128 + lane * 4;
      if (j < j_end) {
        uint2 hi, lo;
        split4_f16(x[t].x * inv, x[t].y * inv, x[t].z * inv, x[t].w * inv, hi, lo);
        *reinterpret_cast<uint2*>(p + j) = hi;
        *reinterpret_cast<uint2*>(p + L + j) = lo;
      }
    }
  }
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_heads_split_f16(const float* src, int64_t ld, int64_t L, int32_t H, int32_t d_k, int32_t a_style,
                                         void* hi, void* lo, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(src && hi && (a_style || lo), GNNLM_E_ARG, "gnnlm_heads_split_f16: null pointer");
  GNNLM_CHECK_ARG(L > 0 && H > 0 && d_k > 0 && d_k % 4 == 0 && ld % 4 == 0 && (uintptr_t)src % 16 == 0, GNNLM_E_SHAPE,
                  "gnnlm_heads_split_f16: d_k and ld must be multiples of 4");
  const int64_t n = L * H * d_k / 4;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  heads_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld, L, H, d_k, a_style, (__half*)hi, (__half*)lo);
  GNNLM_LAUNCH_CHECK("gnnlm_heads_split_f16");
  return 0;
}

extern "C" int32_t gnnlm_heads_transpose_split_f16(const float* src, int64_t ld, int64_t L, int32_t H, int32_t d_k, void* hi,
                                                   void* lo, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(src && hi && lo, GNNLM_E_ARG, "gnnlm_heads_transpose_split_f16: null pointer");
  GNNLM_CHECK_ARG(L > 0 && H > 0 && d_k > 0, GNNLM_E_SHAPE, "gnnlm_heads_transpose_split_f16: bad sizes");
  dim3 grid((unsigned)ceil_div(L, 32), (unsigned)ceil_div(d_k, 32), (unsigned)H);
  heads_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld, L, H, d_k, (__half*)hi, (__half*)lo);
  GNNLM_LAUNCH_CHECK("gnnlm_heads_transpose_split_f16");
  return 0;
}

extern "C" int32_t gnnlm_causal_softmax_split(const float* S, int64_t L, int64_t intra_ctx, int32_t H, int64_t k_tile, void* P,
                                              gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(S && P, GNNLM_E_ARG, "gnnlm_causal_softmax_split: null pointer");
  GNNLM_CHECK_ARG(L > 0 && L % 4 == 0 && H > 0 && k_tile >= 0 && (uintptr_t)S % 16 == 0 && (uintptr_t)P % 8 == 0, GNNLM_E_SHAPE,
                  "gnnlm_causal_softmax_split: L must be a multiple of 4 and S / P 16 B / 8 B aligned");
  int64_t blocks = ceil_div((int64_t)H * L, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  static const bool two_pass = [] { const char* e = getenv("GNNLM_SOFTMAX_TWO_PASS"); return e && e[0] == '1'; }();   // A/B switch
  if (!two_pass && L <= 1024)
    causal_softmax_reg_kernel<8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, L, intra_ctx, H, k_tile, (__half*)P);
  else if (!two_pass && L <= 2048)
    causal_softmax_reg_kernel<16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, L, intra_ctx, H, k_tile, (__half*)P);
  else if (!two_pass && L <= 3072)
    causal_softmax_reg_kernel<24><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, L, intra_ctx, H, k_tile, (__half*)P);
  else
    causal_softmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, L, intra_ctx, H, k_tile, (__half*)P);
  GNNLM_LAUNCH_CHECK("gnnlm_causal_softmax_split");
  return 0;
}
