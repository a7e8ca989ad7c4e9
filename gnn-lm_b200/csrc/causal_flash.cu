// ('tgt','intra','tgt') attention as ONE flash kernel on the tensor cores at fp32 parity.
//
// HGTLayer.forward (reference: fairseq/models/hgt.py:350-358) over the causal edge list of auto_regressive_edges
// (fairseq/data/token_block_dataset.py:586-594: u -> v for u <= v, v - u < max_context when given): per block of L tokens and
// head,  out[v] = sum_{u <= v} softmax_u(<Q[v], K'[u]>) V'[u]  (relation_att / relation_msg / relation_pri / sqrt(d_k) are
// folded into the K' / V' projection weights).  Replaces six launches per layer of the GEMM form (two operand splits, one
// transposed split, S = Q K'^T into an [H, L, L] fp32 buffer, a softmax pass that re-reads it, P V') with one kernel that never
// writes S or P: 1.0 ms per Wiki103 block and three layers -> see DESIGN.md.
//
// Arithmetic = the three-pass fp16 split the projections use, on mma.sync.m16n8k16 (fp32 accumulate):
//   S = Q_hi K_hi^T + Q_hi K_lo^T + Q_lo K_hi^T,   O += P_hi V_hi + P_hi V_lo + P_lo V_hi   (dropped lo x lo terms: 2^-22 relative)
// K' / V' arrive as split fp16 (hi | lo column halves, written by the projection's epilogue), Q as fp32 (split once per CTA
// into A fragments that stay in registers); P = exp2(s - m) is scaled by 2^10 before its split so that its lo half stays a
// normal fp16 number, and the scale leaves with the final 1 / l.
//
// CTA = 4 warps x 16 query rows (64 rows of one head); key / value tiles of 32 rows (K_hi, K_lo, V_hi, V_lo: 32 KB + padding)
// through a 3-deep cp.async ring; rows padded by 16 B so that ldmatrix (K: plain, V: .trans) is bank-conflict free.  CTAs are
// enumerated heaviest first (the last query tile of a block reads L keys, the first 64).
#include "common.cuh"
#include "mma_sync.cuh"

namespace gnnlm {

constexpr int CF_WARPS = 4, CF_THREADS = CF_WARPS * 32;
constexpr int CF_BQ = 16 * CF_WARPS;          // query rows per CTA
constexpr int CF_BK = 32;                     // key rows per tile
constexpr int CF_STAGES = 3;

__device__ __forceinline__ float cf_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DK>
__global__ void __launch_bounds__(CF_THREADS, 2)
    causal_flash_mma_kernel(const float* __restrict__ q, int64_t ldq, const __half* __restrict__ k, const __half* __restrict__ v,
                            int64_t ldkv, int64_t lo_off, int L, int ctx, int H, float* __restrict__ out, int64_t ldo,
                            __half* __restrict__ out_split, int64_t ldos, int64_t os_lo, float out_scale, int accumulate) {
  constexpr int KS = DK / 16;                 // k-steps of the score product
  constexpr int NT = DK / 8;                  // n-tiles of the output
  constexpr int RS = DK * 2 + 16;             // padded row (bytes)
  constexpr int ARR = CF_BK * RS;             // one of the four arrays of a stage
  constexpr int STAGE = 4 * ARR;
  constexpr int CH = DK / 8;                  // 16 B chunks per row
  constexpr float L2E = 1.4426950408889634f;
  extern __shared__ __align__(128) char cf_smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int n_qt = (L + CF_BQ - 1) / CF_BQ;
  const int job = blockIdx.x;
  const int qt = n_qt - 1 - job / H;                       // heaviest query tiles first
  const int h = job % H;
  const int64_t row_base = (int64_t)blockIdx.y * L;        // first token of this block
  const int q0 = qt * CF_BQ;
  const int qw = q0 + warp * 16;                           // first query row of this warp
  const int r0 = qw + g, r1 = qw + g + 8;                  // this thread's two rows

  // ---- Q fragments (fp32 -> fp16 hi / lo), pre-multiplied by log2(e): the softmax runs in the base-2 domain
  uint32_t qh[KS][4], ql[KS][4];
  {
    const float* q0p = q + (row_base + r0) * ldq + h * DK + 2 * tq;
    const float* q1p = q + (row_base + r1) * ldq + h * DK + 2 * tq;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
      if (r0 < L) { a0 = __ldg(reinterpret_cast<const float2*>(q0p + ks * 16)); a2 = __ldg(reinterpret_cast<const float2*>(q0p + ks * 16 + 8)); }
      if (r1 < L) { a1 = __ldg(reinterpret_cast<const float2*>(q1p + ks * 16)); a3 = __ldg(reinterpret_cast<const float2*>(q1p + ks * 16 + 8)); }
      split2_f16(a0.x * L2E, a0.y * L2E, qh[ks][0], ql[ks][0]);
      split2_f16(a1.x * L2E, a1.y * L2E, qh[ks][1], ql[ks][1]);
      split2_f16(a2.x * L2E, a2.y * L2E, qh[ks][2], ql[ks][2]);
      split2_f16(a3.x * L2E, a3.y * L2E, qh[ks][3], ql[ks][3]);
    }
  }

  // ---- key range of this CTA: keys u with u <= v (and v - u < ctx) for some row v of the tile
  const int q_last = (q0 + CF_BQ < L ? q0 + CF_BQ : L) - 1;
  const int k_first = ctx > 0 ? (q0 - ctx + 1 > 0 ? q0 - ctx + 1 : 0) : 0;
  const int t_begin = k_first / CF_BK, t_end = q_last / CF_BK + 1;        // tiles [t_begin, t_end)

  const __half* k_head = k + row_base * ldkv + h * DK;
  const __half* v_head = v + row_base * ldkv + h * DK;
  auto issue = [&](int t, int st) {                        // exactly one commit per call
    if (t < t_end) {
      char* dst = cf_smem + st * STAGE;
      const int kv0 = t * CF_BK;
#pragma unroll
      for (int i = 0; i < 4 * CF_BK * CH / CF_THREADS; ++i) {
        const int c = tid + i * CF_THREADS;
        const int arr = c / (CF_BK * CH), r = (c / CH) % CF_BK, ch = c % CH;
        char* d = dst + arr * ARR + r * RS + ch * 16;
        if (kv0 + r < L) {
          const __half* src = (arr < 2 ? k_head : v_head) + (int64_t)(kv0 + r) * ldkv + ((arr & 1) ? lo_off : 0) + ch * 8;
          cp_async16(d, src);
        } else {
          *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    cp_async_commit();
  };

  float o[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;        // running max (base 2) / partial denominators of rows r0, r1

  const int mi = lane >> 3, r8 = lane & 7;
  const int k_off = ((mi >> 1) * 8 + r8) * RS + (mi & 1) * 16;     // K: matrices (keys 0-7 | 8-15) x (k 0-7 | 8-15)
  const int v_off = ((mi & 1) * 8 + r8) * RS + (mi >> 1) * 16;     // V (.trans): (keys 0-7 | 8-15) x (dv 0-7 | 8-15)

  issue(t_begin, 0);
  issue(t_begin + 1, 1);
  int stage = 0;
  for (int t = t_begin; t < t_end; ++t) {
    asm volatile("cp.async.wait_group %0;" ::"n"(CF_STAGES - 2) : "memory");
    __syncthreads();                                       // tile t visible; everyone is past tile t - 1
    issue(t + 2, stage == 0 ? CF_STAGES - 1 : stage - 1);
    const char* tile = cf_smem + stage * STAGE;
    stage = stage + 1 == CF_STAGES ? 0 : stage + 1;
    const int kv0 = t * CF_BK;
    if (kv0 > qw + 15) continue;                           // above the diagonal for every row of this warp (warp-uniform)
    if (ctx > 0 && kv0 + CF_BK - 1 <= qw - ctx) continue;  // left of the context window for every row of this warp

    // ---- S = Q K^T (base-2 logits): 16 rows x 32 keys
    float s[CF_BK / 8][4];
#pragma unroll
    for (int nt = 0; nt < CF_BK / 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    static_assert(CF_BK == 32, "the score loop below issues the four n-tiles of a 32-key tile pass by pass");
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      // pass-major order: four independent accumulators between two MMAs on the same one (mma.sync latency ~22 clk, issue 8 clk)
      uint32_t bh[2][4], bl[2][4];
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        ldsm_x4(bh[np], tile + np * 16 * RS + ks * 32 + k_off);
        ldsm_x4(bl[np], tile + ARR + np * 16 * RS + ks * 32 + k_off);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(s[nt], ql[ks], bh[nt >> 1][(nt & 1) * 2], bh[nt >> 1][(nt & 1) * 2 + 1]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(s[nt], qh[ks], bl[nt >> 1][(nt & 1) * 2], bl[nt >> 1][(nt & 1) * 2 + 1]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(s[nt], qh[ks], bh[nt >> 1][(nt & 1) * 2], bh[nt >> 1][(nt & 1) * 2 + 1]);
    }
    // ---- mask (diagonal / context-window tiles only)
    if (kv0 + CF_BK - 1 > qw || (ctx > 0 && kv0 <= qw + 15 - ctx)) {
#pragma unroll
      for (int nt = 0; nt < CF_BK / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int u = kv0 + nt * 8 + 2 * tq + (e & 1), vrow = (e & 2) ? r1 : r0;
          if (u > vrow || (ctx > 0 && vrow - u >= ctx)) s[nt][e] = -INFINITY;
        }
      }
    }
    // ---- online softmax of the two rows
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < CF_BK / 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = mx0 > -INFINITY ? cf_exp2(m0 - mx0) : 1.f;        // m = -inf: nothing accumulated yet, o = l = 0
    const float c1 = mx1 > -INFINITY ? cf_exp2(m1 - mx1) : 1.f;
    const float b0 = mx0 > -INFINITY ? mx0 : 0.f, b1 = mx1 > -INFINITY ? mx1 : 0.f;
    m0 = mx0;
    m1 = mx1;
    uint32_t ph[CF_BK / 16][4], pl[CF_BK / 16][4];
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < CF_BK / 8; ++nt) {
      const float p0 = cf_exp2(s[nt][0] - b0), p1 = cf_exp2(s[nt][1] - b0);      // exp2(-inf) = 0 for masked keys
      const float p2 = cf_exp2(s[nt][2] - b1), p3 = cf_exp2(s[nt][3] - b1);
      sum0 += p0 + p1;
      sum1 += p2 + p3;
      // A fragment of P for k-step nt / 2: a0 / a1 from the even n-tile (rows g / g + 8), a2 / a3 from the odd one
      split2_f16(p0 * 1024.f, p1 * 1024.f, ph[nt >> 1][(nt & 1) * 2], pl[nt >> 1][(nt & 1) * 2]);
      split2_f16(p2 * 1024.f, p3 * 1024.f, ph[nt >> 1][(nt & 1) * 2 + 1], pl[nt >> 1][(nt & 1) * 2 + 1]);
    }
    l0 = l0 * c0 + sum0;
    l1 = l1 * c1 + sum1;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1;
    }
    // ---- O += P V
    const char* vt = tile + 2 * ARR;
#pragma unroll
    for (int j = 0; j < CF_BK / 16; ++j) {
#pragma unroll
      for (int nq = 0; nq < NT / 4; ++nq) {              // four n-tiles at a time, pass-major
        uint32_t bh[2][4], bl[2][4];
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          ldsm_x4_t(bh[np], vt + j * 16 * RS + (nq * 2 + np) * 32 + v_off);
          ldsm_x4_t(bl[np], vt + ARR + j * 16 * RS + (nq * 2 + np) * 32 + v_off);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816(o[nq * 4 + nt], pl[j], bh[nt >> 1][(nt & 1) * 2], bh[nt >> 1][(nt & 1) * 2 + 1]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816(o[nq * 4 + nt], ph[j], bl[nt >> 1][(nt & 1) * 2], bl[nt >> 1][(nt & 1) * 2 + 1]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816(o[nq * 4 + nt], ph[j], bh[nt >> 1][(nt & 1) * 2], bh[nt >> 1][(nt & 1) * 2 + 1]);
      }
    }
  }
  cp_async_wait_all();

  // ---- normalise and store: row r (+)= out_scale * O / l
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? out_scale / (1024.f * l0) : 0.f, i1 = l1 > 0.f ? out_scale / (1024.f * l1) : 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = half ? r1 : r0;
    if (r >= L) continue;
    const float inv = half ? i1 : i0;
    float* op = out + (row_base + r) * ldo + h * DK + 2 * tq;
    __half* sp = out_split ? out_split + (row_base + r) * ldos + h * DK + 2 * tq : nullptr;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float2 y = make_float2(o[nt][2 * half] * inv, o[nt][2 * half + 1] * inv);
      if (accumulate) {
        const float2 prev = *reinterpret_cast<const float2*>(op + nt * 8);
        y.x += prev.x;
        y.y += prev.y;
      }
      if (sp) {
        uint32_t hi, lo;
        split2_f16(fminf(fmaxf(y.x, -65504.f), 65504.f), fminf(fmaxf(y.y, -65504.f), 65504.f), hi, lo);
        *reinterpret_cast<uint32_t*>(sp + nt * 8) = hi;
        *reinterpret_cast<uint32_t*>(sp + os_lo + nt * 8) = lo;
      } else {
        *reinterpret_cast<float2*>(op + nt * 8) = y;
      }
    }
  }
}

template <int DK>
static int32_t launch_causal_flash(const float* q, int64_t ldq, const __half* k, const __half* v, int64_t ldkv, int64_t lo_off,
                                   int64_t B, int64_t L, int64_t ctx, int32_t H, float* out, int64_t ldo, __half* out_split,
                                   int64_t ldos, int64_t os_lo, float out_scale, int32_t accumulate, cudaStream_t st) {
  constexpr int smem = CF_STAGES * 4 * CF_BK * (DK * 2 + 16);
  static bool attr_set = false;
  if (!attr_set) {
    GNNLM_CUDA(cudaFuncSetAttribute(causal_flash_mma_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int64_t n_qt = ceil_div(L, CF_BQ);
  dim3 grid((unsigned)(n_qt * H), (unsigned)B);
  causal_flash_mma_kernel<DK><<<grid, CF_THREADS, smem, st>>>(q, ldq, k, v, ldkv, lo_off, (int)L, (int)ctx, H, out, ldo, out_split, ldos,
                                                               os_lo, out_scale, accumulate);
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_causal_flash");
  return 0;
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_hgt_causal_flash(const float* q, int64_t ldq, const void* k_split, const void* v_split, int64_t ldkv,
                                          int64_t lo_off, int64_t B, int64_t L, int64_t intra_ctx, int32_t H, int32_t d_k, float* out,
                                          int64_t ldo, void* out_split, int64_t ldos, int64_t os_lo, float out_scale,
                                          int32_t accumulate, cudaStream_t stream) {
  GNNLM_CHECK_ARG(q && k_split && v_split && (out || (out_split && !accumulate)), GNNLM_E_ARG, "gnnlm_hgt_causal_flash: null pointer");
  GNNLM_CHECK_ARG(B >= 0 && L > 0 && L < (1 << 30) && H > 0 && B < 65536, GNNLM_E_SHAPE, "gnnlm_hgt_causal_flash: bad sizes");
  GNNLM_CHECK_ARG(d_k == 64 || d_k == 128, GNNLM_E_UNSUPPORTED, "gnnlm_hgt_causal_flash: d_k must be 64 or 128 (got %d)", d_k);
  GNNLM_CHECK_ARG(ldq >= (int64_t)H * d_k && ldq % 2 == 0 && (uintptr_t)q % 8 == 0, GNNLM_E_SHAPE,
                  "gnnlm_hgt_causal_flash: q rows must be 8 B aligned");
  GNNLM_CHECK_ARG(ldkv % 8 == 0 && lo_off % 8 == 0 && (uintptr_t)k_split % 16 == 0 && (uintptr_t)v_split % 16 == 0, GNNLM_E_SHAPE,
                  "gnnlm_hgt_causal_flash: K' / V' rows and their lo halves must be 16 B aligned");
  GNNLM_CHECK_ARG(!out || (ldo >= (int64_t)H * d_k && ldo % 2 == 0 && (uintptr_t)out % 8 == 0), GNNLM_E_SHAPE,
                  "gnnlm_hgt_causal_flash: out rows must be 8 B aligned");
  GNNLM_CHECK_ARG(!out_split || (ldos % 2 == 0 && os_lo % 2 == 0 && (uintptr_t)out_split % 4 == 0), GNNLM_E_SHAPE,
                  "gnnlm_hgt_causal_flash: split output rows must be 4 B aligned");
  if (B == 0) return 0;
  const __half* k = reinterpret_cast<const __half*>(k_split);
  const __half* v = reinterpret_cast<const __half*>(v_split);
  __half* os = reinterpret_cast<__half*>(out_split);
  if (d_k == 128)
    return launch_causal_flash<128>(q, ldq, k, v, ldkv, lo_off, B, L, intra_ctx, H, out, ldo, os, ldos, os_lo, out_scale, accumulate, stream);
  return launch_causal_flash<64>(q, ldq, k, v, ldkv, lo_off, B, L, intra_ctx, H, out, ldo, os, ldos, os_lo, out_scale, accumulate, stream);
}
