// ('tgt','intra','tgt') attention as a tcgen05 flash kernel: S = Q K'^T and P V' on the 5th-generation tensor cores with
// accumulators in TMEM, operands staged by TMA, the online softmax in registers between them.  S and P never reach HBM.
//
// Same mathematics and the same three-pass fp16 split as causal_flash.cu (the mma.sync form: hgt.py:350-358 over the causal edges
// of token_block_dataset.py:586-594), on the tensor path that is four times wider (measured: mma.sync peaks at 557 TFLOP/s,
// profiles/r2_mma_sync_peak.log).  d_k = 128.
//
// One CTA per (block, head, 128-query tile), heaviest tiles first; key tiles of 64:
//   warp 0     TMA producer: Q once (hi / lo, two 64-wide k-chunks each, SWIZZLE_128B), then per key tile K' hi / lo (64 x 128) and
//              V'^T hi / lo (128 x 64: V' pre-transposed per head so that it is a K-major B operand), two stages
//   warp 1     MMA issuer (one thread): S[j % 2] = Q K_j^T -- 8 k-steps x 3 passes of tcgen05.mma.kind::f16, M = 128, N = 64 -- issued
//              one tile ahead of the softmax; O[j % 2] = P_j V_j -- 4 k-steps x 3 passes, N = 128 -- when P_j is in shared memory
//   warps 2-5  softmax: thread = query row = TMEM lane.  tcgen05.ld S (64 columns), mask, running max / sum in the base-2 domain,
//              P = 2^7 exp2(s - m_ref) split into fp16 hi / lo and written with tcgen05.st into one of two P buffers IN TMEM (two keys
//              per 32-bit column): the PV product takes P as its TMEM A-operand, so P costs no shared memory, is double-buffered
//              (the softmax of tile j + 1 runs under the PV product of tile j), and the MMA reads only V'^T from shared memory.
//              (The first builds passed P through shared memory in the SWIZZLE_128B layout: one buffer, 224 KB.)  O accumulates in TMEM
//              across the key tiles; it is rescaled (tcgen05.ld -> multiply -> tcgen05.st, per warp) only when a row's maximum
//              has grown by more than 2^8 over the reference the row's P values are scaled by (lazy rescale: P <= 2^8, times the
//              2^7 split scale it stays inside fp16) -- on the first tiles of a row, practically never afterwards
// TMEM: S 3 x 64 columns + O 128 + P 2 x 64 = 448.  Shared memory: Q 64 KB + 2 x (K 32 KB + V^T 32 KB) = 192 KB.
// First build (O partial per tile folded into registers by the softmax threads): 0.119 ms at L = 3072; the softmax warps (one per
// scheduler, ~1200 dependent instructions per tile, half of them fp16 <-> fp32 conversions and the fold) paced it.
#include "gemm_tc_common.cuh"

namespace gnnlm {

namespace ft {

using namespace tc;

constexpr int BQ = 128, BKV = 64, DK = 128;
constexpr int Q_CHUNK = BQ * 128;                      // 16 KB: 128 rows x 64 halves
constexpr int K_CHUNK = BKV * 128;                     // 8 KB
constexpr int VT_TILE = DK * 128;                      // 16 KB: 128 dv rows x 64 keys
constexpr int P_TILE = BQ * 128;                       // 16 KB: 128 rows x 64 keys
constexpr int STAGE = 4 * K_CHUNK + 2 * VT_TILE;       // 64 KB
constexpr int SMEM = 4 * Q_CHUNK + 2 * STAGE;          // P lives in TMEM
constexpr int THREADS = 224;                          // warp 0: Q + K' producer, 1: MMA issuer, 2-5: softmax, 6: V'^T producer
constexpr uint32_t S_COL = 0, O_COL = 256, P_COL = 384, TMEM_ALLOC = 512;   // S: 3 x 64 columns, O: 128, P: 2 x (32 hi + 32 lo)
constexpr float P_SCALE = 128.f, RESCALE_AT = 8.f;   // P = 2^7 exp2(s - m_ref) <= 2^15 as long as the maximum stays within 2^8 of m_ref

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A operand from TMEM (lane = row, one 32-bit column = two consecutive k elements), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Maps {
  CUtensorMap q, k, vt;        // q: box 64 x 128 rows, k: box 64 x 64 rows over the split [T, 4d] matrix; vt: box 64 keys x 128 rows
};

__global__ void __launch_bounds__(THREADS, 1)
    causal_flash_tc_kernel(const __grid_constant__ Maps maps, int L, int ctx, int H, int64_t d, int64_t vt_lo_rows,
                           float* __restrict__ out, int64_t ldo, __half* __restrict__ out_split, int64_t ldos, int64_t os_lo,
                           float out_scale, int accumulate, long long* __restrict__ dbg) {
  constexpr float L2E = 1.4426950408889634f;
  constexpr uint32_t IDESC_QK = (1u << 4) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
  constexpr uint32_t IDESC_PV = (1u << 4) | ((uint32_t)(DK >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
  extern __shared__ uint8_t ft_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                  // hi c0 | hi c1 | lo c0 | lo c1
  uint8_t* sKV = smem + 4 * Q_CHUNK;                   // per stage: K hi c0 | K hi c1 | K lo c0 | K lo c1 | Vt hi | Vt lo
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[3], o_done[2], p_full;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (L + BQ - 1) / BQ;
  const int qt = n_qt - 1 - (int)(blockIdx.x / H);     // heaviest query tiles first
  const int h = (int)(blockIdx.x % H);
  const int b = blockIdx.y;
  const int q0 = qt * BQ;
  const int q_last = (q0 + BQ < L ? q0 + BQ : L) - 1;
  const int k_first = ctx > 0 ? (q0 - ctx + 1 > 0 ? q0 - ctx + 1 : 0) : 0;
  const int t_begin = k_first / BKV, t_end = q_last / BKV + 1;
  const int n_t = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.vt) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 3; ++s) mbar_init(&s_full[s], 1);
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
    mbar_init(&p_full, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_ALLOC)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {                                   // ===================== TMA producer
      const int row_q = b * L + q0;
      const int cq = h * DK;
      mbar_expect_tx(&q_full, 4 * Q_CHUNK);
      tma_load_2d(sQ, &maps.q, cq, row_q, &q_full);
      tma_load_2d(sQ + Q_CHUNK, &maps.q, cq + 64, row_q, &q_full);
      tma_load_2d(sQ + 2 * Q_CHUNK, &maps.q, (int)(2 * d) + cq, row_q, &q_full);
      tma_load_2d(sQ + 3 * Q_CHUNK, &maps.q, (int)(2 * d) + cq + 64, row_q, &q_full);
      const int ck = (int)d + h * DK, ckl = (int)(3 * d) + h * DK;
      // K' and V'^T travel through separate rings: a K' stage is free as soon as its score product has retired (one tile before
      // the PV product that frees the V'^T stage), so the load of K'_{j+2} starts two tile periods before it is needed
      for (int j = 0; j < n_t; ++j) {
        const int st = j & 1;
        mbar_wait(&k_empty[st], (uint32_t)(((j >> 1) & 1) ^ 1));
        uint8_t* s = sKV + st * STAGE;
        const int row_k = b * L + (t_begin + j) * BKV;
        mbar_expect_tx(&k_full[st], 4 * K_CHUNK);
        tma_load_2d(s, &maps.k, ck, row_k, &k_full[st]);
        tma_load_2d(s + K_CHUNK, &maps.k, ck + 64, row_k, &k_full[st]);
        tma_load_2d(s + 2 * K_CHUNK, &maps.k, ckl, row_k, &k_full[st]);
        tma_load_2d(s + 3 * K_CHUNK, &maps.k, ckl + 64, row_k, &k_full[st]);
      }
    }
  } else if (warp == 6) {
    if (lane == 0) {                                   // ===================== V'^T producer
      const int vt_row = (b * H + h) * DK;
      for (int j = 0; j < n_t; ++j) {
        const int st = j & 1;
        mbar_wait(&v_empty[st], (uint32_t)(((j >> 1) & 1) ^ 1));
        uint8_t* s = sKV + st * STAGE + 4 * K_CHUNK;
        const int kv0 = (t_begin + j) * BKV;
        mbar_expect_tx(&v_full[st], 2 * VT_TILE);
        tma_load_2d(s, &maps.vt, kv0, vt_row, &v_full[st]);
        tma_load_2d(s + VT_TILE, &maps.vt, kv0, (int)vt_lo_rows + vt_row, &v_full[st]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                   // ===================== MMA issuer
      const uint64_t dqh0 = make_desc(smem_u32(sQ)), dqh1 = make_desc(smem_u32(sQ + Q_CHUNK));
      const uint64_t dql0 = make_desc(smem_u32(sQ + 2 * Q_CHUNK)), dql1 = make_desc(smem_u32(sQ + 3 * Q_CHUNK));
      long long t_k = 0, t_p = 0, t_v = 0;
      const long long t_begin_clk = dbg ? clock64() : 0;
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        const long long c0 = dbg ? clock64() : 0;
        mbar_wait(&k_full[st], (uint32_t)((j >> 1) & 1));
        if (dbg) t_k += clock64() - c0;
        tc_fence_after();
        const uint8_t* s = sKV + st * STAGE;
        const uint64_t dkh0 = make_desc(smem_u32(s)), dkh1 = make_desc(smem_u32(s + K_CHUNK));
        const uint64_t dkl0 = make_desc(smem_u32(s + 2 * K_CHUNK)), dkl1 = make_desc(smem_u32(s + 3 * K_CHUNK));
        const uint32_t d_tmem = tmem_base + S_COL + (uint32_t)(j % 3) * BKV;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                // 16 halves = 32 B of k per instruction
          const uint64_t off = (uint64_t)((ks & 3) * 2);
          const uint64_t qh = (ks < 4 ? dqh0 : dqh1) + off, ql = (ks < 4 ? dql0 : dql1) + off;
          const uint64_t kh = (ks < 4 ? dkh0 : dkh1) + off, kl = (ks < 4 ? dkl0 : dkl1) + off;
          umma_f16(d_tmem, qh, kl, IDESC_QK, ks > 0 ? 1u : 0u);
          umma_f16(d_tmem, ql, kh, IDESC_QK, 1u);
          umma_f16(d_tmem, qh, kh, IDESC_QK, 1u);
        }
        tc_commit(&k_empty[st]);                        // K'_j consumed
        tc_commit(&s_full[j % 3]);
      };
      mbar_wait(&q_full, 0);
      tc_fence_after();
      // tensor-pipe order ... PV_{j-1}, QK_{j+1}, PV_j, QK_{j+2} ...: the score product of tile j + 2 is issued right AFTER the PV
      // product of tile j, so that it runs under the softmax of tile j + 1 and PV_{j+1} never queues behind it
      issue_qk(0);
      if (n_t > 1) issue_qk(1);
      for (int j = 0; j < n_t; ++j) {
        const int st = j & 1;
        const long long c1 = dbg ? clock64() : 0;
        mbar_wait(&p_full, (uint32_t)(j & 1));          // P_j in shared memory (S_j consumed, O rescaled if it had to be)
        const long long c2 = dbg ? clock64() : 0;
        mbar_wait(&v_full[st], (uint32_t)((j >> 1) & 1));
        if (dbg) { t_p += c2 - c1; t_v += clock64() - c2; }
        tc_fence_after();
        const uint8_t* s = sKV + st * STAGE;
        const uint64_t dvh = make_desc(smem_u32(s + 4 * K_CHUNK)), dvl = make_desc(smem_u32(s + 4 * K_CHUNK + VT_TILE));
        const uint32_t d_tmem = tmem_base + O_COL;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t off = (uint64_t)(ks * 2);
          const uint32_t pa = tmem_base + P_COL + (uint32_t)st * 64 + (uint32_t)ks * 8;      // 16 keys = 8 packed columns
          umma_f16_ts(d_tmem, pa, dvl + off, IDESC_PV, (j | ks) > 0 ? 1u : 0u);          // O accumulates across the key tiles
          umma_f16_ts(d_tmem, pa + 32, dvh + off, IDESC_PV, 1u);
          umma_f16_ts(d_tmem, pa, dvh + off, IDESC_PV, 1u);
        }
        tc_commit(&v_empty[st]);                        // V_j consumed
        tc_commit(&o_done[st]);                         // P_j consumed (its buffer is free), O includes tile j
        if (j + 2 < n_t) issue_qk(j + 2);
      }
      if (dbg && blockIdx.x < 8 && blockIdx.y == 0) {   // the eight heaviest CTAs (timing experiments: GNNLM_FLASH_DEBUG=1)
        dbg[blockIdx.x * 8 + 0] = clock64() - t_begin_clk;
        dbg[blockIdx.x * 8 + 1] = t_k;
        dbg[blockIdx.x * 8 + 2] = t_p;
        dbg[blockIdx.x * 8 + 3] = t_v;
        dbg[blockIdx.x * 8 + 4] = n_t;
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ===================== softmax: thread = query row
    const int lq = warp & 3;                            // TMEM lane quarter of this warp
    const int r = lq * 32 + lane;                       // row inside the tile
    const int v = q0 + r;                               // query position inside the block
    const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
    const uint32_t o_addr = tmem_base + lane_addr + O_COL;
    float m_ref = -INFINITY, l_run = 0.f;               // reference maximum (base 2) the row's P and O are scaled by; denominator
    long long t_s = 0, t_o = 0;
    int n_resc = 0;
    for (int j = 0; j < n_t; ++j) {
      const int st = j & 1;
      const int kv0 = (t_begin + j) * BKV;
      const long long c0 = dbg ? clock64() : 0;
      mbar_wait(&s_full[j % 3], (uint32_t)((j / 3) & 1));
      if (dbg) t_s += clock64() - c0;
      tc_fence_after();
      float s[BKV];
      tmem_ld32_nowait(tmem_base + lane_addr + S_COL + (uint32_t)(j % 3) * BKV, s);
      tmem_ld32_nowait(tmem_base + lane_addr + S_COL + (uint32_t)(j % 3) * BKV + 32, s + 32);
      tmem_wait_ld();
      const bool need_mask = kv0 + BKV - 1 > q0 || (ctx > 0 && kv0 <= q0 + BQ - 1 - ctx);      // CTA-uniform
      float mx = -INFINITY;
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < BKV; ++c) {
          const int u = kv0 + c;
          if (u > v || (ctx > 0 && v - u >= ctx)) s[c] = -INFINITY;
          mx = fmaxf(mx, s[c]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < BKV; ++c) mx = fmaxf(mx, s[c]);
      }
      mx *= L2E;
      // P buffer j & 1 is free once the PV product of tile j - 2 has retired
      if (j > 1) {
        const long long c1 = dbg ? clock64() : 0;
        mbar_wait(&o_done[st], (uint32_t)(((j - 2) >> 1) & 1));
        if (dbg) t_o += clock64() - c1;
      }
      // lazy rescale: move the reference only when the maximum has outgrown it by 2^8 (always on a row's first finite tile)
      const bool grow = mx > m_ref + RESCALE_AT;         // m_ref = -inf: true for any finite mx
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_ref, mx);
        const float corr = m_new > -INFINITY ? fast_exp2(m_ref - m_new) : 1.f;       // 0 when nothing was accumulated under -inf
        m_ref = m_new;
        l_run *= corr;
        ++n_resc;
        if (j > 0) {                                     // O must be complete up to tile j - 1 before it is rescaled
          mbar_wait(&o_done[st ^ 1], (uint32_t)(((j - 1) >> 1) & 1));
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < DK; c0 += 32) {
            float o[32];
            tmem_ld32_nowait(o_addr + c0, o);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] *= corr;
            tmem_st32(o_addr + c0, o);
          }
          tmem_wait_st();
        }
      }
      const float mb = m_ref > -INFINITY ? m_ref : 0.f;
      float sum = 0.f;
      const uint32_t p_addr = tmem_base + lane_addr + P_COL + (uint32_t)st * 64;
      tc_fence_after();
#pragma unroll
      for (int half_ = 0; half_ < 2; ++half_) {           // 32 keys at a time: 16 packed hi + 16 packed lo columns
        float ph[16], pl[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int c = half_ * 32 + 2 * e;
          const float p0 = fast_exp2(fmaf(s[c], L2E, -mb)), p1 = fast_exp2(fmaf(s[c + 1], L2E, -mb));      // exp2(-inf) = 0
          sum += p0 + p1;
          // hi = the top 11 significant bits (exact in fp16, no conversion back needed), lo = the exact fp32 remainder
          const float x0 = p0 * P_SCALE, x1 = p1 * P_SCALE;
          const float h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
          const __half2 hh = __floats2half2_rn(h0, h1), ll = __floats2half2_rn(x0 - h0, x1 - h1);
          ph[e] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&hh));
          pl[e] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&ll));
        }
        tmem_st16(p_addr + half_ * 16, ph);
        tmem_st16(p_addr + 32 + half_ * 16, pl);
      }
      tmem_wait_st();
      l_run += sum;
      tc_fence_before();
      mbar_arrive(&p_full);
    }
    if (dbg && blockIdx.x < 8 && blockIdx.y == 0 && warp == 2 && lane == 0) {
      dbg[blockIdx.x * 8 + 5] = t_s;
      dbg[blockIdx.x * 8 + 6] = t_o;
      dbg[blockIdx.x * 8 + 7] = n_resc;
    }
    // the row's output: O / l
    mbar_wait(&o_done[(n_t - 1) & 1], (uint32_t)(((n_t - 1) >> 1) & 1));
    tc_fence_after();
    float acc[DK];
#pragma unroll
    for (int c0 = 0; c0 < DK; c0 += 32) tmem_ld32_nowait(o_addr + c0, acc + c0);
    tmem_wait_ld();
    if (v < L) {
      const float inv = l_run > 0.f ? out_scale / (P_SCALE * l_run) : 0.f;
      const int64_t row = (int64_t)b * L + v;
      float* op = out + row * ldo + h * DK;
      if (out_split) {
        __half* sp = out_split + row * ldos + h * DK;
#pragma unroll
        for (int c = 0; c < DK; c += 4) {
          float4 y = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
          if (accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(op + c);
            y.x += p.x; y.y += p.y; y.z += p.z; y.w += p.w;
          }
          uint2 hi, lo;
          split4_f16(y.x, y.y, y.z, y.w, hi, lo);
          *reinterpret_cast<uint2*>(sp + c) = hi;
          *reinterpret_cast<uint2*>(sp + os_lo + c) = lo;
        }
      } else {
#pragma unroll
        for (int c = 0; c < DK; c += 4) {
          float4 y = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
          if (accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(op + c);
            y.x += p.x; y.y += p.y; y.z += p.z; y.w += p.w;
          }
          *reinterpret_cast<float4*>(op + c) = y;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_ALLOC) : "memory");
  }
}

// fp16 [rows, cols] row-major (row stride ld elements), boxes of 64 columns x box_rows, SWIZZLE_128B
static int make_map_sw128(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return (int)encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace ft

int32_t gemm_tc_supported();

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int32_t gnnlm_hgt_causal_flash_tc(const void* qk_split, int64_t ldqk, int64_t d, const void* vt, int64_t ldvt, int64_t B,
                                             int64_t L, int64_t intra_ctx, int32_t H, int32_t d_k, float* out, int64_t ldo,
                                             void* out_split, int64_t ldos, int64_t os_lo, float out_scale, int32_t accumulate,
                                             cudaStream_t stream) {
  GNNLM_CHECK_ARG(qk_split && vt && (out || (out_split && !accumulate)), GNNLM_E_ARG, "gnnlm_hgt_causal_flash_tc: null pointer");
  GNNLM_CHECK_ARG(gemm_tc_supported(), GNNLM_E_UNSUPPORTED, "gnnlm_hgt_causal_flash_tc: needs an sm_100 device and driver TMA support");
  GNNLM_CHECK_ARG(d_k == 128 && H > 0 && d == (int64_t)H * d_k, GNNLM_E_UNSUPPORTED, "gnnlm_hgt_causal_flash_tc: d_k must be 128, d = H d_k");
  GNNLM_CHECK_ARG(B >= 0 && B < 65536 && L > 0 && B * L < (1ll << 31), GNNLM_E_SHAPE, "gnnlm_hgt_causal_flash_tc: bad sizes");
  GNNLM_CHECK_ARG(ldqk >= 4 * d && ldqk % 8 == 0 && ldvt >= L && ldvt % 8 == 0 && (uintptr_t)qk_split % 16 == 0 && (uintptr_t)vt % 16 == 0,
                  GNNLM_E_SHAPE, "gnnlm_hgt_causal_flash_tc: TMA needs 16 B aligned bases and row strides (ldqk=%lld ldvt=%lld)",
                  (long long)ldqk, (long long)ldvt);
  GNNLM_CHECK_ARG(!out || (ldo % 4 == 0 && (uintptr_t)out % 16 == 0), GNNLM_E_SHAPE, "gnnlm_hgt_causal_flash_tc: out rows must be 16 B aligned");
  GNNLM_CHECK_ARG(!out_split || (ldos % 4 == 0 && os_lo % 4 == 0 && (uintptr_t)out_split % 8 == 0), GNNLM_E_SHAPE,
                  "gnnlm_hgt_causal_flash_tc: split output rows must be 8 B aligned");
  if (B == 0) return 0;
  ft::Maps maps;
  const int64_t T = B * L, vt_rows = B * H * d_k;
  int r = ft::make_map_sw128(&maps.q, qk_split, T, 4 * d, ldqk, ft::BQ);
  if (!r) r = ft::make_map_sw128(&maps.k, qk_split, T, 4 * d, ldqk, ft::BKV);
  if (!r) r = ft::make_map_sw128(&maps.vt, vt, 2 * vt_rows, L, ldvt, ft::DK);
  GNNLM_CHECK_ARG(r == 0, GNNLM_E_ARG, "gnnlm_hgt_causal_flash_tc: cuTensorMapEncodeTiled failed (%d)", r);
  const int smem = ft::SMEM + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    GNNLM_CUDA(cudaFuncSetAttribute(ft::causal_flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((unsigned)(ceil_div(L, ft::BQ) * H), (unsigned)B);
  static int dbg_on = -1;
  if (dbg_on < 0) { const char* e = getenv("GNNLM_FLASH_DEBUG"); dbg_on = e ? atoi(e) : 0; }       // timing experiments only
  long long* dbg = nullptr;
  if (dbg_on) {
    static long long* buf = nullptr;
    if (!buf) GNNLM_CUDA(cudaMalloc(&buf, 64 * sizeof(long long)));
    GNNLM_CUDA(cudaMemsetAsync(buf, 0, 64 * sizeof(long long), stream));
    dbg = buf;
  }
  ft::causal_flash_tc_kernel<<<grid, ft::THREADS, smem, stream>>>(maps, (int)L, (int)intra_ctx, H, d, vt_rows, out, ldo,
                                                                 reinterpret_cast<__half*>(out_split), ldos, os_lo, out_scale, accumulate,
                                                                 dbg);
  GNNLM_LAUNCH_CHECK("gnnlm_hgt_causal_flash_tc");
  if (dbg) {
    long long host[64];
    GNNLM_CUDA(cudaStreamSynchronize(stream));
    GNNLM_CUDA(cudaMemcpy(host, dbg, sizeof(host), cudaMemcpyDeviceToHost));
    fprintf(stderr, "flash_tc CTA 0: %lld key tiles, issuer total %lld cycles (%.0f per tile): wait K' %lld, wait P %lld, wait V' %lld | "
            "softmax warp: wait S %lld, wait O %lld, rescales %lld\n", host[4], host[0], (double)host[0] / (double)(host[4] ? host[4] : 1),
            host[1], host[2], host[3], host[5], host[6], host[7]);
  }
  return 0;
}
