// Attention backward, head_dim 64, software-pipelined variant of attention_bwd.cu (same math, same
// transposed formulation, same GQA handling; see that file's header).
//
// What the r01 clock trace of the first version showed (tools/attn_trace.py): one (kv tile, q tile)
// iteration took ~7000 cycles against 1280 cycles of tensor work, because everything was a serial
// chain on the softmax threads: softmax 2400 -> wait for dV/dK/dQ MMAs 1500 -> dQ drain 2100 (row-per-lane
// red.global = 32 L2 transactions per warp instruction).  This version
//   * double-buffers the P^T / dS^T operand tiles, and issues S,dP(it+1) BEFORE dV,dK,dQ(it), so the
//     softmax threads of iteration it+1 run while the tensor core finishes iteration it;
//   * keeps P^T in TENSOR MEMORY (packed bf16, 64 columns, written with tcgen05.st) and feeds dV += P^T dO
//     with it as a TMEM A operand: 64 KB less shared-memory traffic per iteration, and the 64 KB of smem the
//     two P^T tiles occupied now hold a third Q/dO stage and a dedicated dQ staging tile;
//   * gives the dQ drain its own warpgroup (warps 10-13): dQ(it) leaves through a swizzled fp32 staging
//     tile as coalesced red.global.add.v4.f32 -- every warp instruction covers two full 256-byte rows.  A stand-alone
//     warpgroup sustains ~20 B/clk/SM this way (tools/microbench/red_pattern.cu: 1600 cycles per tile
//     with all SMs active, 5.7 TB/s chip-wide, L2-resident accumulator), inside the softmax's 2400;
//   * reads LSE / delta as 128-bit shared loads, prefetched one iteration ahead.
// r02: the softmax is split into two phases so that it never waits for the tensor core with the SAME 512 TMEM
// columns: phase A reads S^T and produces P^T (exp2; packed bf16 kept in registers and stored to TMEM for dV),
// phase B reads dP^T and produces dS^T = P^T o (dP^T - delta).  S^T(it+1) is issued the moment phase A(it) has the
// scores in registers (barrier s_free) -- it completes under phase B(it) -- and dP^T(it+1) is issued right after
// phase B(it) (behind dV(it), which frees P^T) -- it completes under phase A(it+1).  The r01 version issued both
// at the end of the softmax and the softmax threads idled ~1350 of every ~3600 cycles waiting for them.
// Tried and dropped (r01 traces): computing S^T/dP^T in two 64-query halves so that half A of tile it+1
// is issued mid-softmax -- the extra operand re-reads slowed the softmax's own shared-memory stores more
// than the overlap gained.  Shared-memory bandwidth is the co-bottleneck here: MMA operand reads
// (~208 KB) + P^T/dS^T stores (64 KB) + dQ staging (64 KB) per iteration = ~2600 cycles at 128 B/clk.
// smem: K 16 + V 16 + Q 3x16 + dO 3x16 + dS^T 2x32 + dQ staging 32 = 224 KB (+2 KB lse/delta + barriers).
// TMEM: S^T 128 | dP^T 128 | dV 64 | dK 64 | dQ 64 | P^T 64 (packed bf16) columns.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace b200 {

namespace {

constexpr int B64_THREADS = 448;  // warp 0 TMA, warp 1 MMA, warps 2-9 two softmax warpgroups, 10-13 dQ drain
constexpr int BT = 128;
constexpr int D = 64;
constexpr int TILE = 128 * D * 2;       // 16 KB
constexpr int PT_BYTES = 128 * 128 * 2;  // 32 KB
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + TILE;
constexpr int QS = 3;                       // Q / dO stages
constexpr int OFF_Q = OFF_V + TILE;
constexpr int OFF_DO = OFF_Q + QS * TILE;
constexpr int OFF_STG = OFF_DO + QS * TILE;  // fp32 dQ staging tile of the drain warpgroup
constexpr int OFF_DS = OFF_STG + PT_BYTES;   // 2 buffers
constexpr int OFF_LSE = OFF_DS + 2 * PT_BYTES;  // [2 slots][lse 128 | delta 128] floats
constexpr int OFF_BAR = OFF_LSE + 2 * 256 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 192;
constexpr int TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384, TM_PT = 448;

struct Bwd64Args {
  const float* lse;
  const float* delta;
  float* dq_acc;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int B, S, H, Hk;
  int dkv_rh;  // heads per dk/dv token row (>= Hk; 2*Hk when dk and dv share one [B,S,2*Hk*D] buffer)
  float scale;
  int causal;
  long long* trace;  // debug: clock64 stamps of CTA (0,0,0), see tools/attn_trace.py
};

#define T64(slot)                                                                             \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 16) \
      p.trace[it * 16 + (slot)] = clock64();                                                  \
  } while (0)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm volatile("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// 8 packed bf16 (4 words) of key row `row`, query chunk `chunk`, into a K-major 128B-swizzled [128 x 128] tile
__device__ __forceinline__ void st_tile_chunk_packed(uint32_t tile, int row, int chunk, const uint32_t* w) {
  const uint32_t addr = tile + (chunk >> 3) * 16384 + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
               : "memory");
}

__global__ void __launch_bounds__(B64_THREADS, 1)
attn_bwd64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                  const Bwd64Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();  // 128B-swizzled TMA/UMMA tiles need a 1 KB aligned base
  const uint32_t sK = sbase + OFF_K, sV = sbase + OFF_V;
  auto sQ = [&](int st) { return sbase + OFF_Q + st * TILE; };
  auto sDO = [&](int st) { return sbase + OFF_DO + st * TILE; };
  const uint32_t sSTG = sbase + OFF_STG;
  auto sDS = [&](int bf) { return sbase + OFF_DS + bf * PT_BYTES; };
  float* lse_s = reinterpret_cast<float*>(smem + OFF_LSE);
  const uint32_t bar = sbase + OFF_BAR;
  const uint32_t kv_full = bar;
  auto q_full = [&](int s) { return bar + 8u * (16 + s); };
  auto q_empty = [&](int s) { return bar + 8u * (19 + s); };
  const uint32_t pt_free = bar + 8u * 1;
  const uint32_t all_done = bar + 8u * 2;  // single phase: every MMA of this CTA retired
  const uint32_t s_full = bar + 8u * 5;   // S^T(it) complete
  const uint32_t dp_full = bar + 8u * 3;  // dP^T(it) complete
  const uint32_t s_free = bar + 8u * 4;   // phase A(it) has read S^T(it): S^T(it+1) may be issued
  const uint32_t pds_full = bar + 8u * 6;
  const uint32_t dq_full = bar + 8u * 7;
  const uint32_t dq_empty = bar + 8u * 8;
  const uint32_t tmem_slot = bar + 8u * 13;

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int kt = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int G = p.H / p.Hk;
  const int k0 = kt * BT;
  const int n_qt_all = (p.S + BT - 1) / BT;
  const int qt_first = p.causal ? kt : 0;
  const int n_qt = n_qt_all - qt_first;
  const int n_it = G * n_qt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < QS; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    mbar_init(pt_free, 1);
    mbar_init(all_done, 1);
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(s_free, 8);
    mbar_init(pds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------- TMA producer ------------------------------------
      mbar_arrive_expect_tx(kv_full, 2 * TILE);
      tma_load_4d(sK, &tmK, kv_full, 0, hk, k0, b);
      tma_load_4d(sV, &tmV, kv_full, 0, hk, k0, b);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % QS;
        const uint32_t ph = (it / QS) & 1u;
        const int h = hk * G + it / n_qt;
        const int q0 = (qt_first + it % n_qt) * BT;
        mbar_wait(q_empty(st), ph ^ 1u);
        mbar_arrive_expect_tx(q_full(st), 2 * TILE);
        tma_load_4d(sQ(st), &tmQ, q_full(st), 0, h, q0, b);
        tma_load_4d(sDO(st), &tmDO, q_full(st), 0, h, q0, b);
      }
    }
  } else if (warp == 1) {
    // ------- MMA issuer: uniform control flow on the whole warp, one elected lane issues --------
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);  // K-major x K-major
    constexpr uint32_t idesc_kv = make_idesc_bf16(128, D, false, true);    // K-major x MN-major
    constexpr uint32_t idesc_dq = make_idesc_bf16(128, D, true, true);     // MN-major x MN-major
    auto issue_s = [&](int it) {  // S^T = K Q^T  (M = keys, N = queries, K = head dim)
      const int st = it % QS;
      mbar_wait(q_full(st), (it / QS) & 1u);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_bf16_ss(tmem_base + TM_S, make_smem_desc_sw128(sK + kk * 32, 0, 1024),
                       make_smem_desc_sw128(sQ(st) + kk * 32, 0, 1024), idesc_s, kk != 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int it) {  // dP^T = V dO^T; the stage's q_full was already observed by issue_s(it)
      const int st = it % QS;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_bf16_ss(tmem_base + TM_DP, make_smem_desc_sw128(sV + kk * 32, 0, 1024),
                       make_smem_desc_sw128(sDO(st) + kk * 32, 0, 1024), idesc_s, kk != 0);
        umma_commit(dp_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    if (n_it > 0) {
      issue_s(0);
      issue_dp(0);
    }
    for (int it = 0; it < n_it; ++it) {
      const int st = it % QS, bf = it & 1;
      if (lane == 0) T64(8);
      if (it + 1 < n_it) {
        mbar_wait(s_free, it & 1u);  // phase A(it) holds the scores in registers: S^T columns are free
        tc_fence_after_sync();
        issue_s(it + 1);             // completes while phase B(it) runs
      }
      if (lane == 0) T64(9);
      mbar_wait(pds_full, it & 1u);  // phase B(it) done: dP^T columns free, P^T (TMEM) and dS^T[bf] written
      if (lane == 0) T64(10);
      if (it > 0) mbar_wait(dq_empty, (it - 1) & 1u);  // dQ(it-1) copied out of TMEM by the drain warps
      tc_fence_after_sync();
      if (lane == 0) T64(11);
      if (elect_one()) {
        // dV += P^T dO  (M = keys, N = head dim, K = queries): A = P^T straight from TMEM, 8 packed columns
        // per K step; issued FIRST so that P^T is free again before phase A(it+1) wants to store into it
#pragma unroll
        for (int kk = 0; kk < BT / 16; ++kk)
          umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_PT + kk * 8,
                       make_smem_desc_sw128(sDO(st) + kk * 2048, 16384, 1024), idesc_kv, (it | kk) != 0);
        umma_commit(pt_free);
      }
      __syncwarp();
      if (it + 1 < n_it) issue_dp(it + 1);  // completes while phase A(it+1) runs
      if (elect_one()) {
        // dK += dS^T Q
#pragma unroll
        for (int kk = 0; kk < BT / 16; ++kk)
          umma_bf16_ss(tmem_base + TM_DK, make_smem_desc_sw128(sDS(bf) + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                       make_smem_desc_sw128(sQ(st) + kk * 2048, 16384, 1024), idesc_kv, (it | kk) != 0);
        umma_commit(q_empty(st));  // Q / dO stage is last read by dV, dK
        // dQ = dS K  (M = queries, N = head dim, K = keys): dS^T tile read as an MN-major A operand
#pragma unroll
        for (int kk = 0; kk < BT / 16; ++kk)
          umma_bf16_ss(tmem_base + TM_DQ, make_smem_desc_sw128(sDS(bf) + kk * 2048, 16384, 1024),
                       make_smem_desc_sw128(sK + kk * 2048, 16384, 1024), idesc_dq, kk != 0);
        umma_commit(dq_full);
      }
      __syncwarp();
      if (lane == 0) T64(12);
    }
    if (elect_one()) umma_commit(all_done);
    __syncwarp();
  } else if (warp >= 10) {
    // ------------------------------------ dQ drain warpgroup --------------------------------------
    const int qd = warp & 3;
    const int row = qd * 32 + lane;  // query row of the dQ tile = TMEM lane
    const int dt = threadIdx.x - 320;  // 0..127
    const uint32_t t_lane = tmem_base + (uint32_t(qd * 32) << 16);
    for (int it = 0; it < n_it; ++it) {
      const int h = hk * G + it / n_qt;
      const int q0 = (qt_first + it % n_qt) * BT;
      mbar_wait(dq_full, it & 1u);  // every MMA of `it` retired
      tc_fence_after_sync();
      if (dt == 0) T64(13);
      uint32_t v[64];
      tmem_ld_32x32b_x32(t_lane + TM_DQ, v);
      tmem_ld_32x32b_x32(t_lane + TM_DQ + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      // staging: [128 rows][16 chunks of 16 B], chunk ^= row & 15 (conflict-free both ways)
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sSTG + row * 256 + ((ch ^ (row & 15)) << 4)),
                     "r"(v[ch * 4 + 0]), "r"(v[ch * 4 + 1]), "r"(v[ch * 4 + 2]), "r"(v[ch * 4 + 3])
                     : "memory");
      }
      named_bar_sync(2, 128);
      float4 val[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = dt + 128 * i;  // 16-byte chunk id: row = c / 16, chunk = c % 16
        const int r = c >> 4, ch = c & 15;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(val[i].x), "=f"(val[i].y), "=f"(val[i].z), "=f"(val[i].w)
                     : "r"(sSTG + r * 256 + ((ch ^ (r & 15)) << 4)));
      }
      named_bar_sync(2, 128);  // everyone has its chunks in registers: the tile may be overwritten
      float* dst = p.dq_acc + (((long long)b * p.S + q0) * p.H + h) * D;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = dt + 128 * i;
        const int r = c >> 4, ch = c & 15;
        if (q0 + r < p.S) red_add_v4(dst + (long long)r * p.H * D + ch * 4, val[i]);
      }
      if (dt == 0) T64(14);
    }
  } else {
    // ------------------------------- softmax / dK,dV epilogue --------------------------------------
    const int qd = warp & 3;
    const int wg = (warp - 2) >> 2;
    const int row = qd * 32 + lane;  // key row (S^T) or query row (dQ)
    const int key = k0 + row;
    const uint32_t t_lane = tmem_base + (uint32_t(qd * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const int tid = threadIdx.x - 64;  // 0..255
    const int cbase = wg * 64;

    // lse*log2e (threads 0-127) / delta (threads 128-255) of the 128 queries of iteration `it`:
    // fetched into a register at the top of the previous iteration, parked in smem after its softmax
    auto fetch_lse = [&](int it) -> float {
      const int h = hk * G + it / n_qt;
      const int q = (qt_first + it % n_qt) * BT + (tid & 127);
      const long long idx = ((long long)b * p.H + h) * p.S + q;
      const float* src = (tid < 128 ? p.lse : p.delta) + (q < p.S ? idx : 0);
      float val;  // asm: keeps the compiler from scheduling the first use (and its stall) right here
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(val) : "l"(src));
      return val;
    };
    auto put_lse = [&](int it, float val) {
      const int q = (qt_first + it % n_qt) * BT + (tid & 127);
      if (tid < 128)
        val = q < p.S ? val * 1.4426950408889634f : INFINITY;
      else
        val = q < p.S ? val : 0.f;
      lse_s[(it & 1) * 256 + tid] = val;
    };
    if (n_it > 0) put_lse(0, fetch_lse(0));
    named_bar_sync(1, 256);
    for (int it = 0; it < n_it; ++it) {
      const int bf = it & 1;
      const int q0 = (qt_first + it % n_qt) * BT;
      const float* lse2 = lse_s + (it & 1) * 256;
      const float* dlt = lse2 + 128;
      const bool diag = p.causal && q0 == k0;  // tiles are aligned: only the diagonal tile is cut
      const bool need_mask = diag || (k0 + BT > p.S);  // CTA-uniform: off-diagonal full tiles skip all per-score tests
      if (threadIdx.x == 64) T64(0);
      float lse_next = 0.f;
      if (it + 1 < n_it) lse_next = fetch_lse(it + 1);  // latency hides behind this iteration's softmax
      // ---------------- phase A: P^T = exp2(S^T * scale*log2e - lse*log2e), packed bf16 ----------------
      mbar_wait(s_full, it & 1u);
      tc_fence_after_sync();
      if (threadIdx.x == 64) T64(1);
      uint32_t pk[32];  // this thread's 64 probabilities (key row x 64 query columns), kept for phase B
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = cbase + hf * 32;
        uint32_t vs[32];
        tmem_ld_32x32b_x32(t_lane + TM_S + c0, vs);
        tmem_ld_wait();
        if (!need_mask) {
          // packed fp32 pairs: the exponent argument costs one issue slot per TWO scores
          const uint64_t sl2_2 = f2_pack(sl2, sl2);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 l4 = *reinterpret_cast<const float4*>(lse2 + c0 + g * 4);
            float a0, a1, a2, a3;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(vs[g * 4 + 0]), __uint_as_float(vs[g * 4 + 1])), sl2_2,
                             f2_pack(-l4.x, -l4.y)), a0, a1);
            f2_unpack(f2_fma(f2_pack(__uint_as_float(vs[g * 4 + 2]), __uint_as_float(vs[g * 4 + 3])), sl2_2,
                             f2_pack(-l4.z, -l4.w)), a2, a3);
            pk[hf * 16 + g * 2] = pack_bf16x2(ex2f(a0), ex2f(a1));
            pk[hf * 16 + g * 2 + 1] = pack_bf16x2(ex2f(a2), ex2f(a3));
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 l4 = *reinterpret_cast<const float4*>(lse2 + c0 + g * 4);
            const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
            float e[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int qc = c0 + g * 4 + i;
              e[i] = ex2f(fmaf(__uint_as_float(vs[g * 4 + i]), sl2, -ls[i]));
              if ((diag && (q0 + qc < key)) || key >= p.S) e[i] = 0.f;
            }
            pk[hf * 16 + g * 2] = pack_bf16x2(e[0], e[1]);
            pk[hf * 16 + g * 2 + 1] = pack_bf16x2(e[2], e[3]);
          }
        }
        // P^T: queries c0 .. c0+31 of this key row = packed columns c0/2 .. c0/2+15.  The previous tile's
        // dV MMAs must have consumed the old contents first.
        if (hf == 0 && it > 0) {
          mbar_wait(pt_free, (it - 1) & 1u);
          tc_fence_after_sync();
        }
        tmem_st_32x32b_x16(t_lane + TM_PT + (c0 >> 1), pk + hf * 16);
      }
      // lse*log2e of the NEXT iteration's queries goes to the other slot now, and only then is S^T handed back:
      // S^T(it+1) completing (s_full) therefore implies every warp's lse(it+1) is in place, and -- because every
      // warp arrives here after its last read of THIS iteration's lse -- the slot being overwritten (last read in
      // phase A(it-1)) is no longer in use.  No CTA-wide barrier in the loop.
      if (it + 1 < n_it && tid < 128) put_lse(it + 1, lse_next);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      if (threadIdx.x == 64) T64(4);
      // ---------------- phase B: dS^T = P^T o (dP^T - delta)  (softmax scale folded out, see below) -------
      // dS here is P o (dP - delta) WITHOUT the softmax scale: dK and dQ are linear in dS, so the scale is
      // applied once per output element (dK epilogue below, dQ conversion kernel) instead of once per score.
      // P is the bf16-rounded probability that also feeds dV (and that the forward kernel multiplied V with).
      mbar_wait(dp_full, it & 1u);
      tc_fence_after_sync();
      if (threadIdx.x == 64) T64(6);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = cbase + hf * 32;
        uint32_t vd[32];
        tmem_ld_32x32b_x32(t_lane + TM_DP + c0, vd);
        tmem_ld_wait();
        // (dP - delta) in packed fp32, rounded to bf16x2, times the bf16x2 probabilities in ONE bf16x2 multiply:
        // the product is already the packed operand the dK / dQ MMAs read (3 issue slots per pair of scores)
        uint32_t dsp[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 d4 = *reinterpret_cast<const float4*>(dlt + c0 + g * 4);
          float a0, a1, a2, a3;
          f2_unpack(f2_sub(f2_pack(__uint_as_float(vd[g * 4 + 0]), __uint_as_float(vd[g * 4 + 1])), f2_pack(d4.x, d4.y)),
                    a0, a1);
          f2_unpack(f2_sub(f2_pack(__uint_as_float(vd[g * 4 + 2]), __uint_as_float(vd[g * 4 + 3])), f2_pack(d4.z, d4.w)),
                    a2, a3);
          dsp[g * 2] = mul_bf16x2(pk[hf * 16 + g * 2], pack_bf16x2(a0, a1));
          dsp[g * 2 + 1] = mul_bf16x2(pk[hf * 16 + g * 2 + 1], pack_bf16x2(a2, a3));
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) st_tile_chunk_packed(sDS(bf), row, (c0 >> 3) + g, dsp + g * 4);
      }
      // delta of the next iteration's queries: parked BEFORE this warp reports phase B done, so dP^T(it+1)
      // completing (dp_full, issued after all eight arrivals) implies the slot is complete; its previous contents
      // (iteration it-1) were last read before pds_full(it-1), which dp_full(it) -- already observed -- followed
      if (it + 1 < n_it && tid >= 128) put_lse(it + 1, lse_next);
      tmem_st_wait();
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (threadIdx.x == 64) T64(2);

    }
    // the softmax warps do not follow dq_full's phases, so a parity wait on it could alias: dedicated barrier
    mbar_wait(all_done, 0);
    tc_fence_after_sync();

    // dK / dV; column halves per warpgroup
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* out = (which == 0 ? p.dv : p.dk) + (((long long)b * p.S + key) * p.dkv_rh + hk) * D + wg * 32;
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_lane + (which == 0 ? TM_DV : TM_DK) + wg * 32, v);
      tmem_ld_wait();
      if (which == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * p.scale);
      }
      if (key < p.S) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o4;
          o4.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
          o4.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
          o4.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
          o4.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
          stg128(out + g * 8, o4);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int make_bshd_map(CUtensorMap* tm, const void* base, int B, int S, int heads, int D, int box_rows, bool f32);

int launch_attn_bwd64(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                      const CUtensorMap& tmDO, const float* lse, const float* delta, float* dq_acc, void* dk,
                      void* dv, int B, int S, int H, int Hk, int dkv_row_heads, float scale, int causal,
                      cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  Bwd64Args a;
  a.lse = lse;
  a.delta = delta;
  a.dq_acc = dq_acc;
  a.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  a.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  a.B = B;
  a.S = S;
  a.H = H;
  a.Hk = Hk;
  a.dkv_rh = dkv_row_heads;
  a.scale = scale;
  a.causal = causal;
  {
    const char* e = getenv("B200_ATTN_TRACE");
    a.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 16)) : nullptr;
  }
  dim3 grid((S + BT - 1) / BT, Hk, B);
  attn_bwd64_kernel<<<grid, B64_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmDO, a);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace b200
