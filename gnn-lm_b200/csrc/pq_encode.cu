// PQ encode -- the step right before the hot path: replaces TorchPQCodec.encode /
// NumpyPQCodec.encode (reference: knn/pq_wrapper.py:51-68,131-167; used by knn/quantize_features.py:115-152 to
// produce quantized-keys.npy).  The OPQ pre-rotation `x @ A.T (+ b)` is a gnnlm_linear call; this kernel does
//   codes[n, m] = argmin_c ( ||cen[m,c]||^2 - 2 <x[n, m*dsub:(m+1)*dsub], cen[m,c]> )      (first minimum wins)
// A warp owns one subspace m for a tile of rows: each lane keeps its 8 centroids (8 x dsub floats) and their
// norms in registers for the whole tile, so per (row, subspace) the only traffic is the 4*dsub-byte x slice
// (read once overall: the M warps of a row tile cover the row) and one code byte.
#include "common.cuh"

namespace gnnlm {

constexpr int ENC_ROWS = 512;      // rows per work item

template <int DSUB>
__global__ void __launch_bounds__(256) pq_encode_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int M,
                                                        const float* __restrict__ cen, const float* __restrict__ norm2,
                                                        uint8_t* __restrict__ codes) {
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (n + ENC_ROWS - 1) / ENC_ROWS;
  const int64_t items = n_tiles * M;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += warps) {
    const int m = (int)(it % M);
    const int64_t r0 = (it / M) * ENC_ROWS;
    float c[8][DSUB], nm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                      // lane owns centroids lane, lane+32, ... (ascending index)
      const int ci = lane + 32 * i;
      nm[i] = __ldg(norm2 + (size_t)m * 256 + ci);
#pragma unroll
      for (int j = 0; j < DSUB; ++j) c[i][j] = __ldg(cen + ((size_t)m * 256 + ci) * DSUB + j);
    }
    const int64_t r1 = r0 + ENC_ROWS < n ? r0 + ENC_ROWS : n;
    for (int64_t r = r0; r < r1; ++r) {
      float xs[DSUB];
#pragma unroll
      for (int j = 0; j < DSUB; ++j) xs[j] = __ldg(x + r * ldx + (size_t)m * DSUB + j);      // warp-uniform: broadcast
      float best = INFINITY;
      int bi = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < DSUB; ++j) dot = fmaf(xs[j], c[i][j], dot);
        const float dis = nm[i] - 2.f * dot;                                                    // pq_wrapper.py:66,164
        if (dis < best) { best = dis; bi = lane + 32 * i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }                       // argmin: first occurrence
      }
      if (lane == 0) codes[r * M + m] = (uint8_t)bi;
    }
  }
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_pq_encode(const float* x, int64_t ldx, int64_t n, int32_t M, int32_t dsub, const float* centroids,
                                   const float* norm2, uint8_t* codes, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(x && centroids && norm2 && codes, GNNLM_E_ARG, "gnnlm_pq_encode: null pointer");
  GNNLM_CHECK_ARG(n >= 0 && M > 0 && ldx >= (int64_t)M * dsub, GNNLM_E_SHAPE, "gnnlm_pq_encode: bad shape");
  if (n == 0) return 0;
  const int64_t items = ceil_div(n, ENC_ROWS) * M;
  int64_t blocks = ceil_div(items, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dsub) {
    case 1: pq_encode_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, n, M, centroids, norm2, codes); break;
    case 2: pq_encode_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, n, M, centroids, norm2, codes); break;
    case 4: pq_encode_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, n, M, centroids, norm2, codes); break;
    case 8: pq_encode_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, n, M, centroids, norm2, codes); break;
    case 16: pq_encode_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, n, M, centroids, norm2, codes); break;
    default:
      set_error("gnnlm_pq_encode: dsub must be one of 1, 2, 4, 8, 16 (got %d)", dsub);
      return GNNLM_E_UNSUPPORTED;
  }
  GNNLM_LAUNCH_CHECK("gnnlm_pq_encode");
  return 0;
}
