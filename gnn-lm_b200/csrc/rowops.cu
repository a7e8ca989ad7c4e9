// Row utilities between stages: row gather, LayerNorm, dtype conversion, tf32 hi/lo weight split.
// Reference call sites: nn.LayerNorm of HGTLayer.forward (fairseq/models/hgt.py:404-405),
// `precompute_feats[offsets].astype(np.float32)` (fairseq/data/token_block_dataset.py:327-329).
// All are HBM-bound streaming kernels: vectorised 16 B accesses, one warp per row.
#include <type_traits>

#include <cuda_fp8.h>

#include "common.cuh"

namespace gnnlm {

template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const V* __restrict__ src, int64_t ld_src,
                                                          const int32_t* __restrict__ ids, V* __restrict__ dst,
                                                          int64_t ld_dst, int64_t n_cap, const int32_t* __restrict__ n_dev,
                                                          int64_t d_vec) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int64_t r = __ldg(ids + i);
    const V* s = src + r * ld_src;
    V* o = dst + i * ld_dst;
    for (int64_t j = lane; j < d_vec; j += 32) o[j] = __ldg(s + j);
  }
}

// --reinit-nfeat (transformer.py:1046-1048): ntgt features are embeddings of the neighbour TOKENS instead of decoded keys.
// out[i, :] = table[labels[rows[i]], :] -- datastore row -> value (the `neighbor_tokens[offset]` of token_block_dataset.py:371,394)
// -> row of the projected embedding table, one warp per node; an out-of-vocabulary value sets *err and yields zeros.
template <typename LT>
__global__ void __launch_bounds__(256) embed_gather_kernel(const uint4* __restrict__ table, int64_t ld_vec, int64_t vocab,
                                                           const LT* __restrict__ labels, int64_t n_datastore,
                                                           const int64_t* __restrict__ rows, const int32_t* __restrict__ row_ids,
                                                           uint4* __restrict__ dst, int64_t ld_dst_vec, int64_t n_cap,
                                                           const int32_t* __restrict__ n_dev, int64_t d_vec,
                                                           int64_t* __restrict__ labels_out, int32_t* __restrict__ err) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int64_t node = row_ids ? (int64_t)__ldg(row_ids + i) : i;
    const int64_t r = __ldg(rows + node);
    int64_t tok = -1;
    if (r >= 0 && r < n_datastore) tok = (int64_t)__ldg(labels + r);
    if (lane == 0 && labels_out) labels_out[i] = tok;
    const bool ok = tok >= 0 && tok < vocab;
    if (!ok && lane == 0 && err) atomicExch(err, 1);
    const uint4* s = table + (ok ? tok : 0) * ld_vec;
    uint4* o = dst + i * ld_dst_vec;
    for (int64_t j = lane; j < d_vec; j += 32) o[j] = ok ? __ldg(s + j) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// residual operand of the fused add + LayerNorm: mode 0 none, 1 fp32, 2 bf16, 3 split-fp16 (lo half d columns later)
__device__ __forceinline__ float res_at(const void* res, int mode, int64_t ld, int64_t row, int64_t j, int64_t d) {
  if (mode == 1) return __ldg(reinterpret_cast<const float*>(res) + row * ld + j);
  if (mode == 2) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(res)[row * ld + j]);
  const __half* p = reinterpret_cast<const __half*>(res) + row * ld + j;
  return __half2float(p[0]) + __half2float(p[d]);
}
__device__ __forceinline__ float4 res4_at(const void* res, int mode, int64_t ld, int64_t row, int64_t j, int64_t d) {
  if (mode == 1) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(res) + row * ld + j));
  if (mode == 2) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(res) + row * ld + j));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  const __half* p = reinterpret_cast<const __half*>(res) + row * ld + j;
  return join4_f16(__ldg(reinterpret_cast<const uint2*>(p)), __ldg(reinterpret_cast<const uint2*>(p + d)));
}

// one warp per row, two passes over registers-resident data when d <= 32*MAXV*4, else re-read
template <typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t ldx, const void* __restrict__ res,
                                                        int res_mode, int64_t ldres,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, OutT* __restrict__ y, int64_t ldy, int64_t n_cap,
                                                        const int32_t* __restrict__ n_dev, int64_t d) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const float* xr = x + i * ldx;
    auto at = [&](int64_t j) { return res_mode ? xr[j] + res_at(res, res_mode, ldres, i, j, d) : xr[j]; };
    float s = 0.f;
    for (int64_t j = lane; j < d; j += 32) s += at(j);
    const float mean = warp_sum(s) / (float)d;
    float vs = 0.f;
    for (int64_t j = lane; j < d; j += 32) {
      float t = at(j) - mean;
      vs = fmaf(t, t, vs);
    }
    const float rstd = rsqrtf(warp_sum(vs) / (float)d + eps);
    for (int64_t j = lane; j < d; j += 32) {
      float v = (at(j) - mean) * rstd * __ldg(gamma + j) + __ldg(beta + j);
      if constexpr (sizeof(OutT) == 4) y[i * ldy + j] = v;
      else if constexpr (std::is_same<OutT, __half>::value) {          // split-fp16: hi | lo
        v = fminf(fmaxf(v, -65504.f), 65504.f);
        const __half h = __float2half_rn(v);
        y[i * ldy + j] = h;
        y[i * ldy + d + j] = __float2half_rn(v - __half2float(h));
      } else y[i * ldy + j] = __float2bfloat16(v);
    }
  }
}

// vectorised variant: d % 128 == 0, row held in registers (d <= 4096)
template <typename OutT, int NV>
__global__ void __launch_bounds__(256, 4) layernorm_vec_kernel(const float* __restrict__ x, int64_t ldx,
                                                            const void* __restrict__ res, int res_mode, int64_t ldres,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            OutT* __restrict__ y, int64_t ldy, int64_t n_cap,
                                                            const int32_t* __restrict__ n_dev, uint8_t* __restrict__ q8 = nullptr,
                                                            int64_t ldq = 0) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int lane = threadIdx.x & 31;
  constexpr int d = NV * 128;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + i * ldx);
    float4 r[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) r[j] = xr[lane + 32 * j];
    if (res_mode) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 t = res4_at(res, res_mode, ldres, i, (int64_t)(lane + 32 * j) * 4, d);
        r[j].x += t.x; r[j].y += t.y; r[j].z += t.z; r[j].w += t.w;
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) s += r[j].x + r[j].y + r[j].z + r[j].w;
    const float mean = warp_sum(s) * (1.f / d);
    float vs = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      r[j].x -= mean; r[j].y -= mean; r[j].z -= mean; r[j].w -= mean;
      vs += r[j].x * r[j].x + r[j].y * r[j].y + r[j].z * r[j].z + r[j].w * r[j].w;
    }
    const float rstd = rsqrtf(warp_sum(vs) * (1.f / d) + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
      float4 o = make_float4(r[j].x * rstd * g.x + b.x, r[j].y * rstd * g.y + b.y, r[j].z * rstd * g.z + b.z,
                             r[j].w * rstd * g.w + b.w);
      if constexpr (sizeof(OutT) == 4) {
        reinterpret_cast<float4*>(y + i * ldy)[lane + 32 * j] = o;
      } else if constexpr (std::is_same<OutT, __half>::value) {        // split-fp16: hi | lo
        uint2 hi, lo;
        split4_f16(o.x, o.y, o.z, o.w, hi, lo);
        reinterpret_cast<uint2*>(y + i * ldy)[lane + 32 * j] = hi;
        reinterpret_cast<uint2*>(y + i * ldy + d)[lane + 32 * j] = lo;
        if (q8) {                                                     // e4m3 companion (hi8 | lo8) for the FP8 correction MMAs
          uint32_t h8, l8;
          q8_from_split4(hi, lo, h8, l8);
          reinterpret_cast<uint32_t*>(q8 + i * ldq)[lane + 32 * j] = h8;
          reinterpret_cast<uint32_t*>(q8 + i * ldq + d)[lane + 32 * j] = l8;
        }
      } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(o.x, o.y), c = __floats2bfloat162_rn(o.z, o.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&c);
        reinterpret_cast<uint2*>(y + i * ldy)[lane + 32 * j] = u;
      }
    }
  }
}

// Deferred LayerNorm (HGT layer 0 with the OPQ rotation folded, hgt.py:401-405): z' = o + x is the pre-norm sum in the UN-rotated
// basis (z = z' rot^T is what the reference normalises; rot orthonormal).  The statistics of z follow from z':
// mean(z) = <z', u> with u = rot^T 1 / d, var(z) = |z' - mean d u|^2 / d.  This kernel writes z' as a gnnlm_linear_f16f8 operand
// (fp16 hi + e4m3 companion) and (mean, 1 / sqrt(var + eps)) per row; the consumers apply the normalisation as a per-row affine
// (the next layer's K' / V': inside gnnlm_hgt_cluster_attn_hq), so neither the rotation of the residual nor the normalised rows of
// the non-centre nodes are ever materialised.  x: fp16 hi [rows, d] + e4m3 companion [rows, hi8 | lo8] (value = hi + lo8 / 2^10).
template <int NV>
__global__ void __launch_bounds__(256, 3) rowstats_q8_kernel(const float* __restrict__ o, int64_t ldo, const __half* __restrict__ x_hi,
                                                             int64_t ldxh, const uint8_t* __restrict__ x_q8, int64_t ldxq,
                                                             const float* __restrict__ u, float eps, __half* __restrict__ y, int64_t ldy,
                                                             uint8_t* __restrict__ q8, int64_t ldq, float2* __restrict__ stats,
                                                             int64_t n_cap, const int32_t* __restrict__ n_dev) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int lane = threadIdx.x & 31;
  constexpr int d = NV * 128;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4* u4 = reinterpret_cast<const float4*>(u) + lane;          // re-read per pass (L1 hits): keeping u in registers spilled
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    float4 r[NV];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      r[j] = reinterpret_cast<const float4*>(o + i * ldo)[lane + 32 * j];
      const uint2 h = __ldg(reinterpret_cast<const uint2*>(x_hi + i * ldxh) + lane + 32 * j);
      const uint32_t l = __ldg(reinterpret_cast<const uint32_t*>(x_q8 + i * ldxq + d) + lane + 32 * j);
      const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
      const __half2_raw l0r = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(l & 0xffff), __NV_E4M3);
      const __half2_raw l1r = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(l >> 16), __NV_E4M3);
      const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l0r)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l1r));
      r[j].x += fmaf(l0.x, 1.f / 1024.f, h0.x); r[j].y += fmaf(l0.y, 1.f / 1024.f, h0.y);
      r[j].z += fmaf(l1.x, 1.f / 1024.f, h1.x); r[j].w += fmaf(l1.y, 1.f / 1024.f, h1.y);
      const float4 uj = __ldg(u4 + 32 * j);
      dot += r[j].x * uj.x + r[j].y * uj.y + r[j].z * uj.z + r[j].w * uj.w;
    }
    const float mean = warp_sum(dot);
    const float md = mean * d;
    float vs = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 uj = __ldg(u4 + 32 * j);
      const float a = r[j].x - md * uj.x, b = r[j].y - md * uj.y, c = r[j].z - md * uj.z, e = r[j].w - md * uj.w;
      vs += a * a + b * b + c * c + e * e;
    }
    const float rstd = rsqrtf(warp_sum(vs) * (1.f / d) + eps);
    if (lane == 0) stats[i] = make_float2(mean, rstd);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      uint2 hi, lo;
      split4_f16(r[j].x, r[j].y, r[j].z, r[j].w, hi, lo);
      reinterpret_cast<uint2*>(y + i * ldy)[lane + 32 * j] = hi;
      uint32_t h8, l8;
      q8_from_split4(hi, lo, h8, l8);
      reinterpret_cast<uint32_t*>(q8 + i * ldq)[lane + 32 * j] = h8;
      reinterpret_cast<uint32_t*>(q8 + i * ldq + d)[lane + 32 * j] = l8;
    }
  }
}

template <typename S, typename D>
__global__ void __launch_bounds__(256) convert_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v;
    if constexpr (sizeof(S) == 4) v = src[i];
    else v = (float)src[i];
    if constexpr (sizeof(D) == 4) dst[i] = v;
    else dst[i] = (D)v;
  }
}

// [rows, d] fp16/fp32 -> split-fp16 [rows, 2d]; one warp per row, 4 elements per lane per step
template <typename S>
__global__ void __launch_bounds__(256) to_split_kernel(const S* __restrict__ src, int64_t ld_src, __half* __restrict__ dst,
                                                       int64_t ld_dst, int64_t rows_cap, const int32_t* __restrict__ rows_dev,
                                                       int64_t d, float scale) {
  const int64_t rows = live_rows(rows_cap, rows_dev);
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const S* s = src + r * ld_src;
    __half* o = dst + r * ld_dst;
    for (int64_t c = lane * 4; c < d; c += 128) {
      float x[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = c + e < d ? scale * (float)s[c + e] : 0.f;      // scale = 1 (exact) except gnnlm_scale_split_f16
      uint2 hi, lo;
      split4_f16(x[0], x[1], x[2], x[3], hi, lo);
      if (c + 3 < d && (d & 3) == 0 && (ld_dst & 3) == 0) {
        *reinterpret_cast<uint2*>(o + c) = hi;
        *reinterpret_cast<uint2*>(o + d + c) = lo;
      } else {
        const __half* hh = reinterpret_cast<const __half*>(&hi);
        const __half* ll = reinterpret_cast<const __half*>(&lo);
        for (int e = 0; e < 4 && c + e < d; ++e) { o[c + e] = hh[e]; o[d + c + e] = ll[e]; }
      }
    }
  }
}

// y = gelu(x) (exact erf form, torch.nn.functional.gelu's default; hgt.py:506) for fp32 rows, written in an activation format
template <int OMODE>      // 0 = f32, 1 = bf16, 2 = split fp16
__global__ void __launch_bounds__(256) gelu_kernel(const float* __restrict__ src, int64_t ld_src, void* __restrict__ dst,
                                                   int64_t ld_dst, int64_t rows_cap, const int32_t* __restrict__ rows_dev,
                                                   int64_t d) {
  const int64_t rows = live_rows(rows_cap, rows_dev);
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float* s = src + r * ld_src;
    for (int64_t c = lane * 4; c < d; c += 128) {             // d % 4 == 0
      const float4 x4 = *reinterpret_cast<const float4*>(s + c);
      float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = 0.5f * x[e] * (1.f + erff(x[e] * 0.70710678118654752f));
      if constexpr (OMODE == 0) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + r * ld_dst + c) = make_float4(x[0], x[1], x[2], x[3]);
      } else if constexpr (OMODE == 2) {
        uint2 hi, lo;
        split4_f16(x[0], x[1], x[2], x[3], hi, lo);
        __half* o = reinterpret_cast<__half*>(dst) + r * ld_dst;
        *reinterpret_cast<uint2*>(o + c) = hi;
        *reinterpret_cast<uint2*>(o + d + c) = lo;
      } else {
        const __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&a);
        u.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + r * ld_dst + c) = u;
      }
    }
  }
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi,
                                                         float* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = w[i];
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);   // keep 10 explicit mantissa bits
    hi[i] = h;
    lo[i] = x - h;                                                        // exact in fp32
  }
}

__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ w, float scale, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = fminf(fmaxf(w[i] * scale, -65504.f), 65504.f);
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn(x - __half2float(h));
  }
}

static inline unsigned grid_for(int64_t n, int per_block) {
  int64_t b = ceil_div(n, per_block);
  const int64_t cap = 148 * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_gather_rows(const void* src, int64_t ld_src, const int32_t* ids, void* dst, int64_t ld_dst,
                                     int64_t n_cap, const int32_t* n_dev, int64_t d, int32_t dtype, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(src && ids && dst, GNNLM_E_ARG, "gnnlm_gather_rows: null pointer");
  const int64_t es = dtype == GNNLM_F32 ? 4 : 2;
  GNNLM_CHECK_ARG(dtype == GNNLM_F32 || dtype == GNNLM_BF16 || dtype == GNNLM_F16, GNNLM_E_UNSUPPORTED, "gnnlm_gather_rows: dtype");
  if (n_cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t row_bytes = d * es;
  const unsigned g = grid_for(n_cap, 8);
  if (row_bytes % 16 == 0 && (ld_src * es) % 16 == 0 && (ld_dst * es) % 16 == 0 && (uintptr_t)src % 16 == 0 &&
      (uintptr_t)dst % 16 == 0) {
    gather_rows_kernel<uint4><<<g, 256, 0, st>>>((const uint4*)src, ld_src * es / 16, ids, (uint4*)dst, ld_dst * es / 16,
                                                n_cap, n_dev, row_bytes / 16);
  } else if (es == 4) {
    gather_rows_kernel<uint32_t><<<g, 256, 0, st>>>((const uint32_t*)src, ld_src, ids, (uint32_t*)dst, ld_dst, n_cap, n_dev, d);
  } else {
    gather_rows_kernel<uint16_t><<<g, 256, 0, st>>>((const uint16_t*)src, ld_src, ids, (uint16_t*)dst, ld_dst, n_cap, n_dev, d);
  }
  GNNLM_LAUNCH_CHECK("gnnlm_gather_rows");
  return 0;
}

extern "C" int32_t gnnlm_embed_gather(const float* table, int64_t ld_table, int64_t vocab, const void* labels, int32_t label_bytes,
                                      int64_t n_datastore, const int64_t* rows, const int32_t* row_ids, float* dst, int64_t ld_dst,
                                      int64_t n_cap, const int32_t* n_dev, int64_t d, int64_t* labels_out, int32_t* err,
                                      gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(table && labels && rows && dst, GNNLM_E_ARG, "gnnlm_embed_gather: null pointer");
  GNNLM_CHECK_ARG(label_bytes == 2 || label_bytes == 4, GNNLM_E_UNSUPPORTED, "gnnlm_embed_gather: labels must be int16 or int32");
  GNNLM_CHECK_ARG(d > 0 && d % 4 == 0 && ld_table % 4 == 0 && ld_dst % 4 == 0 && (uintptr_t)table % 16 == 0 && (uintptr_t)dst % 16 == 0,
                  GNNLM_E_SHAPE, "gnnlm_embed_gather: d and leading dimensions must be multiples of 4 floats, pointers 16 B aligned");
  GNNLM_CHECK_ARG(vocab > 0 && n_datastore > 0 && n_cap >= 0, GNNLM_E_SHAPE, "gnnlm_embed_gather: sizes");
  if (n_cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(n_cap, 8);
  if (label_bytes == 2)
    embed_gather_kernel<int16_t><<<g, 256, 0, st>>>((const uint4*)table, ld_table / 4, vocab, (const int16_t*)labels, n_datastore, rows,
                                                   row_ids, (uint4*)dst, ld_dst / 4, n_cap, n_dev, d / 4, labels_out, err);
  else
    embed_gather_kernel<int32_t><<<g, 256, 0, st>>>((const uint4*)table, ld_table / 4, vocab, (const int32_t*)labels, n_datastore, rows,
                                                   row_ids, (uint4*)dst, ld_dst / 4, n_cap, n_dev, d / 4, labels_out, err);
  GNNLM_LAUNCH_CHECK("gnnlm_embed_gather");
  return 0;
}

extern "C" int32_t gnnlm_layernorm(const float* x, int64_t ldx, const void* residual, int32_t r_dtype, int64_t ldr,
                                   const float* gamma, const float* beta, float eps, void* y, int32_t out_dtype, int64_t ldy,
                                   int64_t n_cap, const int32_t* n_dev, int64_t d, gnnlm_stream_t stream) {
  return gnnlm_layernorm_q8(x, ldx, residual, r_dtype, ldr, gamma, beta, eps, y, out_dtype, ldy, nullptr, 0, n_cap, n_dev, d, stream);
}

extern "C" int32_t gnnlm_layernorm_q8(const float* x, int64_t ldx, const void* residual, int32_t r_dtype, int64_t ldr,
                                      const float* gamma, const float* beta, float eps, void* y, int32_t out_dtype, int64_t ldy,
                                      void* q8v, int64_t ldq, int64_t n_cap, const int32_t* n_dev, int64_t d,
                                      gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(x && gamma && beta && y, GNNLM_E_ARG, "gnnlm_layernorm: null pointer");
  uint8_t* q8 = reinterpret_cast<uint8_t*>(q8v);
  GNNLM_CHECK_ARG(!q8 || (out_dtype == GNNLM_F16X2 && d % 128 == 0 && d <= 1024 && ldq >= 2 * d && ldq % 4 == 0 && (uintptr_t)q8 % 4 == 0),
                  GNNLM_E_UNSUPPORTED, "gnnlm_layernorm_q8: the e4m3 companion needs a split-fp16 output, d in {128..1024} a multiple of 128, ldq >= 2d");
  const int res_mode = !residual ? 0 : (r_dtype == GNNLM_F32 ? 1 : (r_dtype == GNNLM_BF16 ? 2 : (r_dtype == GNNLM_F16X2 ? 3 : -1)));
  GNNLM_CHECK_ARG(res_mode >= 0, GNNLM_E_UNSUPPORTED, "gnnlm_layernorm: residual dtype");
  GNNLM_CHECK_ARG(!residual || ldr >= d * (res_mode == 3 ? 2 : 1), GNNLM_E_SHAPE, "gnnlm_layernorm: ldr too small");
  const void* res = residual;
  const int64_t ldres = ldr;
  GNNLM_CHECK_ARG(out_dtype == GNNLM_F32 || out_dtype == GNNLM_BF16 || out_dtype == GNNLM_F16X2, GNNLM_E_UNSUPPORTED,
                  "gnnlm_layernorm: out dtype");
  GNNLM_CHECK_ARG(out_dtype != GNNLM_F16X2 || ldy >= 2 * d, GNNLM_E_SHAPE, "gnnlm_layernorm: split output needs ldy >= 2d");
  GNNLM_CHECK_ARG(d > 0, GNNLM_E_SHAPE, "gnnlm_layernorm: d");
  if (n_cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(n_cap, 8);
  const bool vec = d % 128 == 0 && d <= 1024 && ldx % 4 == 0 && ldy % 4 == 0 && (uintptr_t)x % 16 == 0 &&
                   (!residual || ((uintptr_t)residual % 16 == 0 && ldr % 8 == 0)) && (uintptr_t)y % 16 == 0 && (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0;
  GNNLM_CHECK_ARG(!q8 || (vec && (d == 1024 || d == 512 || d == 256 || d == 128)), GNNLM_E_UNSUPPORTED,
                  "gnnlm_layernorm_q8: the e4m3 companion is written by the vectorised kernel only (d in {128, 256, 512, 1024}, 16 B aligned rows)");
#define LN_VEC(NV)                                                                                                      \
  if (out_dtype == GNNLM_F32)                                                                                           \
    layernorm_vec_kernel<float, NV><<<g, 256, 0, st>>>(x, ldx, res, res_mode, ldres, gamma, beta, eps, (float*)y, ldy, n_cap, n_dev);          \
  else if (out_dtype == GNNLM_F16X2)                                                                                    \
    layernorm_vec_kernel<__half, NV><<<g, 256, 0, st>>>(x, ldx, res, res_mode, ldres, gamma, beta, eps, (__half*)y, ldy, n_cap, n_dev, q8, ldq); \
  else                                                                                                                  \
    layernorm_vec_kernel<__nv_bfloat16, NV><<<g, 256, 0, st>>>(x, ldx, res, res_mode, ldres, gamma, beta, eps, (__nv_bfloat16*)y, ldy, n_cap, n_dev);
  if (vec && d == 1024) { LN_VEC(8) }
  else if (vec && d == 512) { LN_VEC(4) }
  else if (vec && d == 256) { LN_VEC(2) }
  else if (vec && d == 128) { LN_VEC(1) }
  else if (out_dtype == GNNLM_F32)
    layernorm_kernel<float><<<g, 256, 0, st>>>(x, ldx, res, res_mode, ldres, gamma, beta, eps, (float*)y, ldy, n_cap, n_dev, d);
  else if (out_dtype == GNNLM_F16X2)
    layernorm_kernel<__half><<<g, 256, 0, st>>>(x, ldx, res, res_mode, ldres, gamma, beta, eps, (__half*)y, ldy, n_cap, n_dev, d);
  else
    layernorm_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(x, ldx, res, res_mode, ldres, gamma, beta, eps, (__nv_bfloat16*)y, ldy, n_cap, n_dev, d);
#undef LN_VEC
  GNNLM_LAUNCH_CHECK("gnnlm_layernorm");
  return 0;
}

extern "C" int32_t gnnlm_rowstats_q8(const float* o, int64_t ldo, const void* x_hi, int64_t ldxh, const void* x_q8, int64_t ldxq,
                                     const float* u, float eps, void* y_hi, int64_t ldy, void* y_q8, int64_t ldq, float* stats,
                                     int64_t n_cap, const int32_t* n_dev, int64_t d, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(o && x_hi && x_q8 && u && y_hi && y_q8 && stats, GNNLM_E_ARG, "gnnlm_rowstats_q8: null pointer");
  GNNLM_CHECK_ARG(d == 128 || d == 256 || d == 512 || d == 1024, GNNLM_E_UNSUPPORTED, "gnnlm_rowstats_q8: d in {128, 256, 512, 1024}");
  GNNLM_CHECK_ARG(ldo % 4 == 0 && ldo >= d && ldxh % 4 == 0 && ldxh >= d && ldxq % 4 == 0 && ldxq >= 2 * d && ldy % 4 == 0 && ldy >= d &&
                      ldq % 4 == 0 && ldq >= 2 * d && (uintptr_t)o % 16 == 0 && (uintptr_t)x_hi % 8 == 0 && (uintptr_t)x_q8 % 4 == 0 &&
                      (uintptr_t)u % 16 == 0 && (uintptr_t)y_hi % 8 == 0 && (uintptr_t)y_q8 % 4 == 0 && (uintptr_t)stats % 8 == 0,
                  GNNLM_E_SHAPE, "gnnlm_rowstats_q8: alignment / leading dimensions");
  if (n_cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(n_cap, 8);
#define RS_VEC(NV)                                                                                                                \
  rowstats_q8_kernel<NV><<<g, 256, 0, st>>>(o, ldo, (const __half*)x_hi, ldxh, (const uint8_t*)x_q8, ldxq, u, eps, (__half*)y_hi, ldy, \
                                           (uint8_t*)y_q8, ldq, reinterpret_cast<float2*>(stats), n_cap, n_dev)
  if (d == 1024) RS_VEC(8);
  else if (d == 512) RS_VEC(4);
  else if (d == 256) RS_VEC(2);
  else RS_VEC(1);
#undef RS_VEC
  GNNLM_LAUNCH_CHECK("gnnlm_rowstats_q8");
  return 0;
}

extern "C" int32_t gnnlm_convert(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n,
                                 gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(src && dst, GNNLM_E_ARG, "gnnlm_convert: null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(n, 1024);
  if (src_dtype == GNNLM_F16 && dst_dtype == GNNLM_F32) convert_kernel<__half, float><<<g, 256, 0, st>>>((const __half*)src, (float*)dst, n);
  else if (src_dtype == GNNLM_F16 && dst_dtype == GNNLM_BF16) convert_kernel<__half, __nv_bfloat16><<<g, 256, 0, st>>>((const __half*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == GNNLM_F32 && dst_dtype == GNNLM_BF16) convert_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == GNNLM_BF16 && dst_dtype == GNNLM_F32) convert_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == GNNLM_F32 && dst_dtype == GNNLM_F16) convert_kernel<float, __half><<<g, 256, 0, st>>>((const float*)src, (__half*)dst, n);
  else if (src_dtype == GNNLM_F32 && dst_dtype == GNNLM_F32) convert_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else {
    set_error("gnnlm_convert: unsupported conversion %d -> %d", src_dtype, dst_dtype);
    return GNNLM_E_UNSUPPORTED;
  }
  GNNLM_LAUNCH_CHECK("gnnlm_convert");
  return 0;
}

extern "C" int32_t gnnlm_split_tf32(const float* w, float* w_hi, float* w_lo, int64_t n, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(w && w_hi && w_lo, GNNLM_E_ARG, "gnnlm_split_tf32: null pointer");
  if (n == 0) return 0;
  split_tf32_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)stream>>>(w, w_hi, w_lo, n);
  GNNLM_LAUNCH_CHECK("gnnlm_split_tf32");
  return 0;
}

extern "C" int32_t gnnlm_split_f16(const float* w, float scale, void* w_hi, void* w_lo, int64_t n, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(w && w_hi && w_lo, GNNLM_E_ARG, "gnnlm_split_f16: null pointer");
  GNNLM_CHECK_ARG(scale > 0.f, GNNLM_E_ARG, "gnnlm_split_f16: scale must be positive");
  if (n == 0) return 0;
  split_f16_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)stream>>>(w, scale, (__half*)w_hi, (__half*)w_lo, n);
  GNNLM_LAUNCH_CHECK("gnnlm_split_f16");
  return 0;
}

extern "C" int32_t gnnlm_to_split_f16(const void* src, int32_t src_dtype, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows,
                                      const int32_t* rows_dev, int64_t d, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(src && dst, GNNLM_E_ARG, "gnnlm_to_split_f16: null pointer");
  GNNLM_CHECK_ARG(src_dtype == GNNLM_F32 || src_dtype == GNNLM_F16, GNNLM_E_UNSUPPORTED, "gnnlm_to_split_f16: source must be F32 or F16");
  GNNLM_CHECK_ARG(d > 0 && ld_src >= d && ld_dst >= 2 * d, GNNLM_E_SHAPE, "gnnlm_to_split_f16: bad shape");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(rows, 8);
  if (src_dtype == GNNLM_F32) to_split_kernel<float><<<g, 256, 0, st>>>((const float*)src, ld_src, (__half*)dst, ld_dst, rows, rows_dev, d, 1.f);
  else to_split_kernel<__half><<<g, 256, 0, st>>>((const __half*)src, ld_src, (__half*)dst, ld_dst, rows, rows_dev, d, 1.f);
  GNNLM_LAUNCH_CHECK("gnnlm_to_split_f16");
  return 0;
}

// split-fp16 of scale * src: a gradient operand brought into the fp16 range by a power of two on its way into the operand format
// (train.py: the product divides the scale out) instead of a scaled fp32 copy first
extern "C" int32_t gnnlm_scale_split_f16(const float* src, int64_t ld_src, float scale, void* dst, int64_t ld_dst, int64_t rows,
                                         const int32_t* rows_dev, int64_t d, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(src && dst, GNNLM_E_ARG, "gnnlm_scale_split_f16: null pointer");
  GNNLM_CHECK_ARG(d > 0 && ld_src >= d && ld_dst >= 2 * d, GNNLM_E_SHAPE, "gnnlm_scale_split_f16: bad shape");
  if (rows == 0) return 0;
  to_split_kernel<float><<<grid_for(rows, 8), 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__half*)dst, ld_dst, rows, rows_dev, d, scale);
  GNNLM_LAUNCH_CHECK("gnnlm_scale_split_f16");
  return 0;
}

extern "C" int32_t gnnlm_gelu(const float* src, int64_t ld_src, void* dst, int32_t dst_dtype, int64_t ld_dst, int64_t rows,
                              const int32_t* rows_dev, int64_t d, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(src && dst, GNNLM_E_ARG, "gnnlm_gelu: null pointer");
  GNNLM_CHECK_ARG(dst_dtype == GNNLM_F32 || dst_dtype == GNNLM_BF16 || dst_dtype == GNNLM_F16X2, GNNLM_E_UNSUPPORTED,
                  "gnnlm_gelu: output must be F32, BF16 or F16X2");
  GNNLM_CHECK_ARG(d > 0 && d % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 && ld_src >= d &&
                      ld_dst >= (dst_dtype == GNNLM_F16X2 ? 2 * d : d) && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0,
                  GNNLM_E_SHAPE, "gnnlm_gelu: d and the leading dimensions must be multiples of 4, pointers 16 B aligned");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(rows, 8);
  if (dst_dtype == GNNLM_F32) gelu_kernel<0><<<g, 256, 0, st>>>(src, ld_src, dst, ld_dst, rows, rows_dev, d);
  else if (dst_dtype == GNNLM_F16X2) gelu_kernel<2><<<g, 256, 0, st>>>(src, ld_src, dst, ld_dst, rows, rows_dev, d);
  else gelu_kernel<1><<<g, 256, 0, st>>>(src, ld_src, dst, ld_dst, rows, rows_dev, d);
  GNNLM_LAUNCH_CHECK("gnnlm_gelu");
  return 0;
}
