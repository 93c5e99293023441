// Attention backward, head_dim 128 (BASELINE C5), software-pipelined like attention_bwd64.cu: same math, same
// transposed formulation (keys on the 128 TMEM lanes), same GQA handling -- see attention_bwd.cu's header.
//
// The first version (attention_bwd.cu, still used as B200_ATTN_BWD_V1=1) ran every (kv tile, q tile) iteration as
// one serial chain -- blocking lse/delta loads + CTA barrier, Q/dO TMA (single stage), S^T/dP^T MMAs, softmax,
// dV/dK/dQ MMAs, dQ drain by the softmax warps -- about 13 000 cycles per iteration on the C5 shape against
// 2 560 cycles of tensor work.  With head dim 128 every resource is full (TMEM: S^T 128 | dP^T 128 | dV 128 | dK 128;
// smem: seven 32 KB tiles), so the pipeline is built from what becomes free EARLY inside an iteration:
//   * two-phase softmax: phase A turns S^T into P^T (packed bf16, kept in registers), phase B turns dP^T into dS^T
//     and only then writes P^T and dS^T to their (single) smem tiles.  S^T's columns are free after phase A, so
//     S^T(it+1) is issued mid-softmax (second Q stage) and is complete before phase A(it+1) begins.
//   * dQ(it) is accumulated in the dP^T columns (free after phase B) and drained by a dedicated warpgroup while the
//     softmax warps already run phase A(it+1); dP^T(it+1) is issued as soon as the drain has copied dQ(it) out of TMEM.
//     The fp32 reductions into the dQ accumulator (64 KB per iteration) must not sit on that chain: issued from
//     registers (row per lane, or coalesced through a staging tile) they kept the drain busy for ~2700 cycles and
//     phase B(it+1) waiting for its operand tiles.  The two operand tiles are idle between their MMAs and phase
//     B(it+1), so the drain parks dQ(it) in them in the TMA 128B-swizzle layout and ONE thread hands the four
//     [128 x 32] fp32 boxes to cp.reduce.async.bulk.tensor (.add): the tiles are free again as soon as the TMA engine
//     has READ them (wait_group.read), the L2 reductions complete in the background.
//   * MMA order after phase B(it): dV (frees dO for the next TMA load), dQ (starts the drain), dK.
//   * lse / delta of the next iteration are prefetched into registers and parked in the other smem slot without a
//     CTA-wide barrier (ordering argument in attention_bwd64.cu).
// smem: K 32 + V 32 + Q 2x32 + dO 32 + P^T 32 + dS^T 32 = 224 KB (+2 KB lse/delta + barriers).
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace b200 {

namespace {

constexpr int B128_THREADS = 448;  // warp 0 TMA, warp 1 MMA, warps 2-9 two softmax warpgroups, 10-13 dQ drain
constexpr int BT = 128;
constexpr int D = 128;
constexpr int TILE = 128 * D * 2;  // 32 KB (two 64-column halves of 16 KB, 128B swizzle each)
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + TILE;
constexpr int OFF_Q = OFF_V + TILE;  // 2 stages
constexpr int OFF_DO = OFF_Q + 2 * TILE;
constexpr int OFF_PT = OFF_DO + TILE;
constexpr int OFF_DS = OFF_PT + TILE;
constexpr int OFF_LSE = OFF_DS + TILE;  // [2 slots][lse 128 | delta 128] floats
constexpr int OFF_BAR = OFF_LSE + 2 * 256 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 192;
constexpr int TM_S = 0, TM_DP = 128, TM_DQ = 128, TM_DV = 256, TM_DK = 384;

struct Bwd128Args {
  const float* lse;
  const float* delta;
  float* dq_acc;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int B, S, H, Hk;
  int dkv_rh;  // heads per dk/dv token row (>= Hk)
  float scale;
  int causal;
  long long* trace;  // debug: clock64 stamps of CTA (0,0,0), see tools/attn_trace.py
};

#define T128(slot)                                                                              \
  do {                                                                                          \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 16) \
      p.trace[it * 16 + (slot)] = clock64();                                                    \
  } while (0)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

__global__ void __launch_bounds__(B128_THREADS, 1)
attn_bwd128_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDQ, const Bwd128Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();  // 128B-swizzled TMA/UMMA tiles need a 1 KB aligned base
  const uint32_t sK = sbase + OFF_K, sV = sbase + OFF_V, sDO = sbase + OFF_DO;
  auto sQ = [&](int st) { return sbase + OFF_Q + st * TILE; };
  const uint32_t sPT = sbase + OFF_PT, sDS = sbase + OFF_DS;
  float* lse_s = reinterpret_cast<float*>(smem + OFF_LSE);
  const uint32_t bar = sbase + OFF_BAR;
  const uint32_t kv_full = bar;
  auto q_full = [&](int s) { return bar + 8u * (1 + s); };
  auto q_empty = [&](int s) { return bar + 8u * (3 + s); };
  const uint32_t do_full = bar + 8u * 5;
  const uint32_t do_empty = bar + 8u * 6;
  const uint32_t s_full = bar + 8u * 7;     // S^T(it) complete
  const uint32_t dp_full = bar + 8u * 8;    // dP^T(it) complete
  const uint32_t s_free = bar + 8u * 9;     // phase A(it) has read S^T(it)
  const uint32_t pds_full = bar + 8u * 10;  // phase B(it) done: P^T, dS^T written, dP^T read
  const uint32_t pds_free = bar + 8u * 11;  // dV, dQ, dK(it) retired: P^T / dS^T tiles may be rewritten
  const uint32_t dq_full = bar + 8u * 12;
  const uint32_t dq_empty = bar + 8u * 13;
  const uint32_t all_done = bar + 8u * 14;
  const uint32_t stg_free = bar + 8u * 16;  // drain(it) no longer uses the P^T tile as its staging buffer
  const uint32_t tmem_slot = bar + 8u * 15;

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
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(s_free, 8);
    mbar_init(pds_full, 8);
    mbar_init(pds_free, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    mbar_init(all_done, 1);
    mbar_init(stg_free, 1);
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
#pragma unroll
      for (int db = 0; db < 2; ++db) {
        tma_load_4d(sK + db * 16384, &tmK, kv_full, db * 64, hk, k0, b);
        tma_load_4d(sV + db * 16384, &tmV, kv_full, db * 64, hk, k0, b);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const int h = hk * G + it / n_qt;
        const int q0 = (qt_first + it % n_qt) * BT;
        mbar_wait(q_empty(st), ((it >> 1) & 1u) ^ 1u);  // dK(it-2) retired
        mbar_arrive_expect_tx(q_full(st), TILE);
#pragma unroll
        for (int db = 0; db < 2; ++db) tma_load_4d(sQ(st) + db * 16384, &tmQ, q_full(st), db * 64, h, q0, b);
        mbar_wait(do_empty, (it & 1u) ^ 1u);  // dV(it-1) retired
        mbar_arrive_expect_tx(do_full, TILE);
#pragma unroll
        for (int db = 0; db < 2; ++db) tma_load_4d(sDO + db * 16384, &tmDO, do_full, db * 64, h, q0, b);
      }
    }
  } else if (warp == 1) {
    // ------- MMA issuer: uniform control flow on the whole warp, one elected lane issues --------
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);  // K-major x K-major
    constexpr uint32_t idesc_kv = make_idesc_bf16(128, D, false, true);    // K-major x MN-major
    constexpr uint32_t idesc_dq = make_idesc_bf16(128, D, true, true);     // MN-major x MN-major
    auto issue_s = [&](int it) {  // S^T = K Q^T  (M = keys, N = queries, K = head dim)
      const int st = it & 1;
      mbar_wait(q_full(st), (it >> 1) & 1u);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_bf16_ss(tmem_base + TM_S, make_smem_desc_sw128(sK + off, 0, 1024),
                       make_smem_desc_sw128(sQ(st) + off, 0, 1024), idesc_s, kk != 0);
        }
        umma_commit(s_full);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int it) {  // dP^T = V dO^T
      mbar_wait(do_full, it & 1u);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_bf16_ss(tmem_base + TM_DP, make_smem_desc_sw128(sV + off, 0, 1024),
                       make_smem_desc_sw128(sDO + off, 0, 1024), idesc_s, kk != 0);
        }
        umma_commit(dp_full);
      }
      __syncwarp();
    };
    // loop-invariant operand descriptors: a K step advances the 14-bit address field (bytes >> 4) by a constant, so
    // each MMA costs one 64-bit add per operand instead of a descriptor rebuild (the issue loop competes for its
    // scheduler with two softmax warps and a drain warp: 24 MMAs took ~1600 cycles to ISSUE in the first version)
    const uint64_t d_pt = make_smem_desc_sw128(sPT, 0, 1024);         // K-major A: +(kk>>2)*16384 + (kk&3)*32 bytes
    const uint64_t d_ds_k = make_smem_desc_sw128(sDS, 0, 1024);       // dS^T as K-major A (dK)
    const uint64_t d_ds_mn = make_smem_desc_sw128(sDS, 16384, 1024);  // dS^T as MN-major A (dQ): + kk*2048 bytes
    const uint64_t d_do_mn = make_smem_desc_sw128(sDO, 16384, 1024);  // dO as MN-major B (dV)
    const uint64_t d_k_mn = make_smem_desc_sw128(sK, 16384, 1024);    // K as MN-major B (dQ)
    const uint64_t d_q_mn[2] = {make_smem_desc_sw128(sQ(0), 16384, 1024), make_smem_desc_sw128(sQ(1), 16384, 1024)};
    mbar_wait(kv_full, 0);
    if (n_it > 0) {
      issue_s(0);
      issue_dp(0);
    }
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      if (lane == 0) T128(8);
      if (it + 1 < n_it) {
        mbar_wait(s_free, it & 1u);  // phase A(it) holds the scores in registers: S^T columns are free
        tc_fence_after_sync();
        issue_s(it + 1);             // completes while phase B(it) runs
      }
      if (lane == 0) T128(9);
      mbar_wait(pds_full, it & 1u);  // phase B(it) done: P^T / dS^T tiles written, dP^T columns free
      tc_fence_after_sync();
      if (lane == 0) T128(10);
      if (elect_one()) {
        // dQ = dS K  (M = queries, N = head dim, K = keys) FIRST: it heads the longest dependent chain (drain ->
        // dP^T(it+1) -> phase B(it+1)).  dS^T tile read as an MN-major A operand; accumulator = the dP^T columns
        // (phase B(it) has consumed them, and dQ(it-1) was drained before dP^T(it) was issued)
#pragma unroll
        for (int kk = 0; kk < BT / 16; ++kk)
          umma_bf16_ss(tmem_base + TM_DQ, d_ds_mn + (uint64_t)(kk * 128), d_k_mn + (uint64_t)(kk * 128), idesc_dq, kk != 0);
        umma_commit(dq_full);
        // dV += P^T dO  (M = keys, N = head dim, K = queries)
#pragma unroll
        for (int kk = 0; kk < BT / 16; ++kk)
          umma_bf16_ss(tmem_base + TM_DV, d_pt + (uint64_t)((kk >> 2) * 1024 + (kk & 3) * 2),
                       d_do_mn + (uint64_t)(kk * 128), idesc_kv, (it | kk) != 0);
        umma_commit(do_empty);  // dO tile free for the next TMA load; P^T tile free for the drain's staging
        // dK += dS^T Q
#pragma unroll
        for (int kk = 0; kk < BT / 16; ++kk)
          umma_bf16_ss(tmem_base + TM_DK, d_ds_k + (uint64_t)((kk >> 2) * 1024 + (kk & 3) * 2),
                       d_q_mn[st] + (uint64_t)(kk * 128), idesc_kv, (it | kk) != 0);
        umma_commit(q_empty(st));
        umma_commit(pds_free);  // covers dQ, dV and dK: both operand tiles may be overwritten
      }
      __syncwarp();
      if (lane == 0) T128(11);
      if (it + 1 < n_it) {
        mbar_wait(dq_empty, it & 1u);  // dQ(it) copied out of the dP^T columns by the drain warps
        tc_fence_after_sync();
        if (lane == 0) T128(12);
        issue_dp(it + 1);              // completes while phase A(it+1) runs
      }
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
      mbar_wait(dq_full, it & 1u);
      tc_fence_after_sync();
      if (dt == 0) T128(13);
      // half 0 of the fp32 dQ tile is staged in the P^T tile (free once dV(it) retired: do_empty), half 1 in the dS^T
      // tile (free once dK(it) retired: pds_free); layout = two [128 rows x 32 fp32] TMA boxes per tile, 128B swizzle
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[64];
        tmem_ld_32x32b_x32(t_lane + TM_DQ + hf * 64, v);
        tmem_ld_32x32b_x32(t_lane + TM_DQ + hf * 64 + 32, v + 32);
        tmem_ld_wait();
        if (hf == 1) {  // all of dQ(it) has left tensor memory: dP^T(it+1) may be issued
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(dq_empty);
        }
        const uint32_t stg = hf == 0 ? sPT : sDS;
        mbar_wait(hf == 0 ? do_empty : pds_free, it & 1u);
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (ch >> 3) * 16384 + row * 128 +
                                                                         (((ch & 7) ^ (row & 7)) << 4)),
                       "r"(v[ch * 4 + 0]), "r"(v[ch * 4 + 1]), "r"(v[ch * 4 + 2]), "r"(v[ch * 4 + 3])
                       : "memory");
        }
        fence_proxy_async_smem();  // the staged values must be visible to the TMA engine
        named_bar_sync(2, 128);
        if (dt == 0) {
          // this half goes out at once (the L2 reduction rate, ~22 B/clk/SM, is what the drain waits for); rows
          // beyond S are clipped by the tensor map; concurrent CTAs (other key tiles) reduce atomically
          tma_reduce_add_4d(&tmDQ, stg, hf * 64, h, q0, b);
          tma_reduce_add_4d(&tmDQ, stg + 16384, hf * 64 + 32, h, q0, b);
          tma_commit_group();
        }
      }
      if (dt == 0) tma_wait_group_read<0>();  // the engine has read both tiles: phase B(it+1) may overwrite them
      if (dt == 0) {
        T128(14);
        mbar_arrive(stg_free);
      }
    }
    if (dt == 0) tma_wait_group<0>();  // every reduction has been performed before the CTA retires
  } else {
    // ------------------------------- softmax / dK,dV epilogue --------------------------------------
    const int qd = warp & 3;
    const int wg = (warp - 2) >> 2;
    const int row = qd * 32 + lane;  // key row
    const int key = k0 + row;
    const uint32_t t_lane = tmem_base + (uint32_t(qd * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const int tid = threadIdx.x - 64;  // 0..255
    const int cbase = wg * 64;

    // lse*log2e (threads 0-127) / delta (threads 128-255) of the 128 queries of iteration `it`
    auto fetch_lse = [&](int it) -> float {
      const int h = hk * G + it / n_qt;
      const int q = (qt_first + it % n_qt) * BT + (tid & 127);
      const long long idx = ((long long)b * p.H + h) * p.S + q;
      const float* src = (tid < 128 ? p.lse : p.delta) + (q < p.S ? idx : 0);
      float val;
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
      const int q0 = (qt_first + it % n_qt) * BT;
      const float* lse2 = lse_s + (it & 1) * 256;
      const float* dlt = lse2 + 128;
      const bool diag = p.causal && q0 == k0;  // tiles are aligned: only the diagonal tile is cut
      const bool need_mask = diag || (k0 + BT > p.S);
      float lse_next = 0.f;
      if (it + 1 < n_it) lse_next = fetch_lse(it + 1);  // latency hides behind this iteration's softmax
      // ---------------- phase A: P^T = exp2(S^T * scale*log2e - lse*log2e), packed bf16 in registers ---------
      if (threadIdx.x == 64) T128(0);
      mbar_wait(s_full, it & 1u);
      tc_fence_after_sync();
      if (threadIdx.x == 64) T128(1);
      uint32_t pk[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = cbase + hf * 32;
        uint32_t vs[32];
        tmem_ld_32x32b_x32(t_lane + TM_S + c0, vs);
        tmem_ld_wait();
        if (!need_mask) {
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
      }
      // next iteration's lse, then hand S^T back (ordering argument: attention_bwd64.cu)
      if (it + 1 < n_it && tid < 128) put_lse(it + 1, lse_next);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      if (threadIdx.x == 64) T128(4);
      // ---------------- phase B: dS^T = P^T o (dP^T - delta); P^T and dS^T to their smem tiles --------------
      mbar_wait(dp_full, it & 1u);
      if (threadIdx.x == 64) T128(5);
      if (it > 0) {
        mbar_wait(pds_free, (it - 1) & 1u);  // dV, dQ, dK(it-1) no longer read the operand tiles
        mbar_wait(stg_free, (it - 1) & 1u);  // and the dQ drain is done with the P^T tile (its staging buffer)
      }
      tc_fence_after_sync();
      if (threadIdx.x == 64) T128(6);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = cbase + hf * 32;
        uint32_t vd[32];
        tmem_ld_32x32b_x32(t_lane + TM_DP + c0, vd);
        tmem_ld_wait();
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
        for (int g = 0; g < 4; ++g) {
          st_tile_chunk_packed(sPT, row, (c0 >> 3) + g, pk + hf * 16 + g * 4);
          st_tile_chunk_packed(sDS, row, (c0 >> 3) + g, dsp + g * 4);
        }
      }
      if (it + 1 < n_it && tid >= 128) put_lse(it + 1, lse_next);  // delta of the next iteration
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (threadIdx.x == 64) T128(2);
    }
    mbar_wait(all_done, 0);
    tc_fence_after_sync();

    // dK / dV; column halves per warpgroup.  dS carried no softmax scale: dK gets it here, dQ in the conversion pass
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* out = (which == 0 ? p.dv : p.dk) + (((long long)b * p.S + key) * p.dkv_rh + hk) * D + wg * 64;
      const float mul = which == 1 ? p.scale : 1.0f;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + (which == 0 ? TM_DV : TM_DK) + wg * 64 + c0, v);
        tmem_ld_wait();
        if (key < p.S) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * mul, __uint_as_float(v[g * 8 + 1]) * mul);
            o4.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * mul, __uint_as_float(v[g * 8 + 3]) * mul);
            o4.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * mul, __uint_as_float(v[g * 8 + 5]) * mul);
            o4.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * mul, __uint_as_float(v[g * 8 + 7]) * mul);
            stg128(out + c0 + g * 8, o4);
          }
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

int launch_attn_bwd128(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                       const CUtensorMap& tmDO, const float* lse, const float* delta, float* dq_acc, void* dk,
                       void* dv, int B, int S, int H, int Hk, int dkv_row_heads, float scale, int causal,
                       cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  Bwd128Args a;
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
  CUtensorMap tmDQ;  // fp32 [B, S, H, 128] accumulator, boxes of [128 rows x 32 floats], 128B swizzle
  if (int rc = make_bshd_map(&tmDQ, dq_acc, B, S, H, D, BT, true)) return rc;
  dim3 grid((S + BT - 1) / BT, Hk, B);
  attn_bwd128_kernel<<<grid, B128_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmDO, tmDQ, a);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace b200
