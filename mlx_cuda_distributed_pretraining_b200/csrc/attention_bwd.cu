// Fused causal / GQA attention backward on tcgen05 (sm_100a): the adjoint the reference obtains
// from MLX autograd of arch/flash_attention.py:123-151 (core/training.py:1580,1650), including
// the sum of dK/dV over the query heads that share a kv head (adjoint of mx.repeat, :107-120).
//
//   delta = rowsum(dO * O)                                  (pre-pass kernel)
//   per (kv tile, kv head, batch) CTA, looping over the group's query heads and query tiles:
//     S^T  = K Q^T          P^T  = exp2(S^T * scale*log2e - LSE*log2e)         (keys on TMEM lanes)
//     dP^T = V dO^T         dS^T = P^T * (dP^T - delta) * scale
//     dV  += P^T dO         dK  += dS^T Q          dQ_tile = dS K  -> fp32 red.global into dq_acc
//   dq_acc (fp32) -> dq (bf16)                               (post-pass kernel)
//
// Working transposed puts keys on the 128 TMEM lanes, so P^T / dS^T rows are written by their
// owning thread straight into K-major swizzled smem tiles that feed the dV / dK MMAs, and the
// same dS^T tile is read as an MN-major A operand for dQ -- no smem transposes anywhere.
// Warp roles: warp 0 TMA, warp 1 MMA issue, warps 2-9 (two warpgroups, half the columns each)
// softmax / dQ drain / dK,dV epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace b200 {

int make_bshd_map(CUtensorMap* tm, const void* base, int B, int S, int heads, int D, int box_rows,
                  bool f32);
int launch_attn_bwd64(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                      const CUtensorMap& tmDO, const float* lse, const float* delta, float* dq_acc, void* dk,
                      void* dv, int B, int S, int H, int Hk, int dkv_row_heads, float scale, int causal,
                      cudaStream_t stream);

int launch_attn_bwd128(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                       const CUtensorMap& tmDO, const float* lse, const float* delta, float* dq_acc, void* dk,
                       void* dv, int B, int S, int H, int Hk, int dkv_row_heads, float scale, int causal,
                       cudaStream_t stream);

namespace {

constexpr int BWD_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9: two softmax warpgroups
constexpr int BT = 128;  // tile size along both queries and keys

template <int D>
struct BwdCfg {
  static constexpr int ST = (D == 64) ? 2 : 1;      // Q/dO stages
  static constexpr bool ALIAS_DQ = (D == 128);      // dQ accumulator reuses the S^T columns
  static constexpr int TILE_BYTES = 128 * D * 2;
  static constexpr int PT_BYTES = 128 * 128 * 2;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE_BYTES;
  static constexpr int OFF_Q = OFF_V + TILE_BYTES;
  static constexpr int OFF_DO = OFF_Q + ST * TILE_BYTES;
  static constexpr int OFF_PT = OFF_DO + ST * TILE_BYTES;
  static constexpr int OFF_DS = OFF_PT + PT_BYTES;
  static constexpr int OFF_LSE = OFF_DS + PT_BYTES;  // 2 slots x (128 lse + 128 delta) floats
  static constexpr int OFF_BAR = OFF_LSE + 2 * 2 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 256 + D;
  static constexpr int TM_DQ = ALIAS_DQ ? 0 : 256 + 2 * D;
  static constexpr int TMEM_COLS = 512;
};

struct BwdArgs {
  const float* lse;    // [B,H,S]
  const float* delta;  // [B,H,S]
  float* dq_acc;       // [B,S,H,D] fp32
  __nv_bfloat16* dk;   // [B,S,Hk,D] rows of dkv_rh heads
  __nv_bfloat16* dv;
  int B, S, H, Hk;
  int dkv_rh;          // heads per dk/dv token row (>= Hk)
  float scale;
  int causal;
  long long* trace;  // optional clock64 trace of CTA 0 (debug; nullptr in production)
};

#define BWD_TRACE(slot)                                                                   \
  do {                                                                                    \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 16) \
      p.trace[it * 16 + (slot)] = clock64();                                              \
  } while (0)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// store 8 bf16 (one 16-byte chunk) of row `row`, logical column chunk `chunk` (0..15) of a
// [128 x 128] K-major tile made of two 64-column blocks with 128-byte swizzle
__device__ __forceinline__ void st_tile_chunk(uint32_t tile, int row, int chunk, const float* f) {
  const uint32_t addr = tile + (chunk >> 3) * 16384 + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(f[0], f[1])),
               "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])),
               "r"(pack_bf16x2(f[6], f[7]))
               : "memory");
}

template <int D>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const BwdArgs p) {
  using Cfg = BwdCfg<D>;
  constexpr int ST = Cfg::ST;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));  // generic pointer to the aligned base
  const uint32_t sK = sbase + Cfg::OFF_K, sV = sbase + Cfg::OFF_V;
  auto sQ = [&](int st) { return sbase + Cfg::OFF_Q + st * Cfg::TILE_BYTES; };
  auto sDO = [&](int st) { return sbase + Cfg::OFF_DO + st * Cfg::TILE_BYTES; };
  const uint32_t sPT = sbase + Cfg::OFF_PT, sDS = sbase + Cfg::OFF_DS;
  float* lse_s = reinterpret_cast<float*>(sgen + Cfg::OFF_LSE);  // [2][2][128]
  const uint32_t bar = sbase + Cfg::OFF_BAR;
  const uint32_t kv_full = bar;
  auto q_full = [&](int s) { return bar + 8u * (1 + s); };
  auto q_empty = [&](int s) { return bar + 8u * (3 + s); };
  const uint32_t s_full = bar + 8u * 5;
  const uint32_t pds_full = bar + 8u * 6;
  const uint32_t dq_full = bar + 8u * 7;
  const uint32_t dq_empty = bar + 8u * 8;
  const uint32_t tmem_slot = bar + 8u * 9;

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;

  const int kt = blockIdx.x;  // kv tile (tile 0 sees every query tile under the causal mask)
  const int hk = blockIdx.y;
  const int b = blockIdx.z;
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
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------- TMA producer ------------------------------------
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
      for (int db = 0; db < D / 64; ++db) {
        tma_load_4d(sK + db * 16384, &tmK, kv_full, db * 64, hk, k0, b);
        tma_load_4d(sV + db * 16384, &tmV, kv_full, db * 64, hk, k0, b);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it % ST;
        const uint32_t ph = (it / ST) & 1u;
        const int h = hk * G + it / n_qt;
        const int q0 = (qt_first + it % n_qt) * BT;
        BWD_TRACE(6);
        mbar_wait(q_empty(st), ph ^ 1u);
        BWD_TRACE(7);
        mbar_arrive_expect_tx(q_full(st), 2 * Cfg::TILE_BYTES);
#pragma unroll
        for (int db = 0; db < D / 64; ++db) {
          tma_load_4d(sQ(st) + db * 16384, &tmQ, q_full(st), db * 64, h, q0, b);
          tma_load_4d(sDO(st) + db * 16384, &tmDO, q_full(st), db * 64, h, q0, b);
        }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer: the whole warp runs the control flow (uniform), -------------
      // ---------------- one elected lane issues tcgen05.mma / commit ---------------------------
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);  // K-major x K-major
      constexpr uint32_t idesc_kv = make_idesc_bf16(128, D, false, true);    // K-major x MN-major
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, D, true, true);     // MN-major x MN-major
      mbar_wait(kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % ST;
        if (lane == 0) BWD_TRACE(8);
        mbar_wait(q_full(st), (it / ST) & 1u);
        if (Cfg::ALIAS_DQ && it > 0) mbar_wait(dq_empty, (it - 1) & 1u);
        tc_fence_after_sync();
        if (lane == 0) BWD_TRACE(9);
        if (elect_one()) {
          // S^T = K Q^T and dP^T = V dO^T   (M = keys, N = queries, K = head dim)
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
            umma_bf16_ss(tmem_base + Cfg::TM_S, make_smem_desc_sw128(sK + off, 0, 1024),
                         make_smem_desc_sw128(sQ(st) + off, 0, 1024), idesc_s, kk != 0);
          }
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
            umma_bf16_ss(tmem_base + Cfg::TM_DP, make_smem_desc_sw128(sV + off, 0, 1024),
                         make_smem_desc_sw128(sDO(st) + off, 0, 1024), idesc_s, kk != 0);
          }
          umma_commit(s_full);
        }
        __syncwarp();
        if (lane == 0) BWD_TRACE(10);

        mbar_wait(pds_full, it & 1u);
        if (lane == 0) BWD_TRACE(11);
        if (!Cfg::ALIAS_DQ && it > 0) mbar_wait(dq_empty, (it - 1) & 1u);
        tc_fence_after_sync();
        if (lane == 0) BWD_TRACE(12);
        if (elect_one()) {
          // dV += P^T dO ; dK += dS^T Q   (M = keys, N = head dim, K = queries)
#pragma unroll
          for (int kk = 0; kk < BT / 16; ++kk) {
            const uint32_t aoff = (kk >> 2) * 16384 + (kk & 3) * 32;
            umma_bf16_ss(tmem_base + Cfg::TM_DV, make_smem_desc_sw128(sPT + aoff, 0, 1024),
                         make_smem_desc_sw128(sDO(st) + kk * 2048, 16384, 1024), idesc_kv,
                         (it | kk) != 0);
          }
#pragma unroll
          for (int kk = 0; kk < BT / 16; ++kk) {
            const uint32_t aoff = (kk >> 2) * 16384 + (kk & 3) * 32;
            umma_bf16_ss(tmem_base + Cfg::TM_DK, make_smem_desc_sw128(sDS + aoff, 0, 1024),
                         make_smem_desc_sw128(sQ(st) + kk * 2048, 16384, 1024), idesc_kv,
                         (it | kk) != 0);
          }
          // dQ = dS K   (M = queries, N = head dim, K = keys): dS^T tile read as an MN-major A
#pragma unroll
          for (int kk = 0; kk < BT / 16; ++kk) {
            umma_bf16_ss(tmem_base + Cfg::TM_DQ, make_smem_desc_sw128(sDS + kk * 2048, 16384, 1024),
                         make_smem_desc_sw128(sK + kk * 2048, 16384, 1024), idesc_dq, kk != 0);
          }
          umma_commit(q_empty(st));
          umma_commit(dq_full);
        }
        __syncwarp();
        if (lane == 0) BWD_TRACE(13);
      }
    }
  } else {
    // --------------------------- softmax / dQ drain / dK,dV epilogue ------------------------
    const int qd = warp & 3;
    const int row = qd * 32 + lane;  // key row inside the tile (S^T) or query row (dQ)
    const int key = k0 + row;
    const uint32_t t_lane = tmem_base + (uint32_t(qd * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const int tid = threadIdx.x - 64;  // 0..255 (warp-major, not TMEM-lane order)
    // Two warpgroups share every tile: both see all 128 key rows (TMEM lane quarter = warp % 4) and
    // each takes half of the columns, so every SM sub-partition holds two softmax warps that hide
    // each other's TMEM / MUFU latency.
    const int wg = (warp - 2) >> 2;

    for (int it = 0; it < n_it; ++it) {
      const int h = hk * G + it / n_qt;
      const int q0 = (qt_first + it % n_qt) * BT;
      float* lse2 = lse_s + (it & 1) * 256;
      float* dlt = lse2 + 128;
      if (threadIdx.x == 64) BWD_TRACE(0);
      {
        const int qi = tid & 127;
        const int q = q0 + qi;
        const long long idx = ((long long)b * p.H + h) * p.S + q;
        if (tid < 128)
          lse2[qi] = q < p.S ? p.lse[idx] * 1.4426950408889634f : INFINITY;
        else
          dlt[qi] = q < p.S ? p.delta[idx] : 0.f;
      }
      named_bar_sync(1, 256);
      if (threadIdx.x == 64) BWD_TRACE(1);
      mbar_wait(s_full, it & 1u);
      tc_fence_after_sync();
      if (threadIdx.x == 64) BWD_TRACE(2);
      const bool diag = p.causal && q0 == k0;  // tiles are aligned: only the diagonal tile is cut
#pragma unroll 1
      for (int c0 = wg * (BT / 2); c0 < (wg + 1) * (BT / 2); c0 += 32) {
        uint32_t vs[32], vd[32];
        tmem_ld_32x32b_x32(t_lane + Cfg::TM_S + c0, vs);
        tmem_ld_32x32b_x32(t_lane + Cfg::TM_DP + c0, vd);
        tmem_ld_wait();
        float pr[32], ds[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int qc = c0 + i;
          float e = ex2(fmaf(__uint_as_float(vs[i]), sl2, -lse2[qc]));
          if ((diag && (q0 + qc < key)) || key >= p.S) e = 0.f;
          pr[i] = e;
          ds[i] = e * (__uint_as_float(vd[i]) - dlt[qc]) * p.scale;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          st_tile_chunk(sPT, row, (c0 >> 3) + g, pr + g * 8);
          st_tile_chunk(sDS, row, (c0 >> 3) + g, ds + g * 8);
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (threadIdx.x == 64) BWD_TRACE(3);

      // drain dQ (TMEM lanes are query rows here) into the fp32 accumulator
      mbar_wait(dq_full, it & 1u);
      tc_fence_after_sync();
      if (threadIdx.x == 64) BWD_TRACE(4);
      {
        const int q = q0 + row;
        float* dst = p.dq_acc + (((long long)b * p.S + q) * p.H + h) * D;
#pragma unroll 1
        for (int c0 = wg * (D / 2); c0 < (wg + 1) * (D / 2); c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_lane + Cfg::TM_DQ + c0, v);
          tmem_ld_wait();
          if (q < p.S) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              red_add_v4(dst + c0 + g * 4, __uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1]),
                         __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      if (threadIdx.x == 64) BWD_TRACE(5);
    }

    // dK / dV: the last dq_full commit covers every MMA issued before it
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* out = (which == 0 ? p.dv : p.dk) + (((long long)b * p.S + key) * p.dkv_rh + hk) * D;
      const uint32_t col = which == 0 ? Cfg::TM_DV : Cfg::TM_DK;
#pragma unroll 1
      for (int c0 = wg * (D / 2); c0 < (wg + 1) * (D / 2); c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + col + c0, v);
        tmem_ld_wait();
        if (key < p.S) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o4;
            if (n_it > 0) {
              o4.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
              o4.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
              o4.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
              o4.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            } else {
              o4 = make_uint4(0, 0, 0, 0);
            }
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
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// delta[b,h,s] = sum_d o[b,s,h,d] * do[b,s,h,d]; LPR = D/8 lanes cooperate on one row.  The same pass zeroes the
// fp32 dQ accumulator (same [B,S,H,D] index space: each thread clears the 8 floats under the 8 bf16 it reads), so the
// backward needs no separate memset launch over 4*B*S*H*D bytes.
template <int D>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                  float* __restrict__ delta, float* __restrict__ dq_acc, int B, int S, int H) {
  constexpr int LPR = D / 8;
  const long long rows = (long long)B * S * H;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = gtid / LPR;
  const int sub = (int)(gtid % LPR);
  float acc = 0.f;
  if (r < rows) {
    const uint4 a = ldg128(o + r * D + sub * 8);
    const uint4 c = ldg128(d_o + r * D + sub * 8);
    float4* z = reinterpret_cast<float4*>(dq_acc + r * D + sub * 8);
    z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t* au = &a.x;
    const uint32_t* cu = &c.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16x2(au[i]), y = unpack_bf16x2(cu[i]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
#pragma unroll
  for (int off = LPR / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (r < rows && sub == 0) {
    const int hh = (int)(r % H);
    const long long bs = r / H;
    const int s = (int)(bs % S);
    const long long bb = bs / S;
    delta[(bb * H + hh) * S + s] = acc;
  }
}

__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long nvec, float mul) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(src + i * 8);
    const float4 c = *reinterpret_cast<const float4*>(src + i * 8 + 4);
    uint4 o4;
    o4.x = pack_bf16x2(a.x * mul, a.y * mul);
    o4.y = pack_bf16x2(a.z * mul, a.w * mul);
    o4.z = pack_bf16x2(c.x * mul, c.y * mul);
    o4.w = pack_bf16x2(c.z * mul, c.w * mul);
    stg128(dst + i * 8, o4);
  }
}

template <int D>
int launch_bwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
               const CUtensorMap& tmDO, const BwdArgs& a, cudaStream_t stream) {
  using Cfg = BwdCfg<D>;
  auto kern = attn_bwd_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((a.S + BT - 1) / BT, a.Hk, a.B);
  kern<<<grid, BWD_THREADS, Cfg::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmDO, a);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

size_t attn_bwd_workspace_bytes(int B, int S, int H, int Hk, int D) {
  (void)Hk;
  return al256((size_t)B * S * H * D * 4) + al256((size_t)B * H * S * 4);
}

int attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
             const float* lse, void* dq, void* dk, void* dv, int B, int S, int H, int Hk, int D,
             float scale, int causal, int dkv_row_heads, void* ws, size_t ws_bytes, cudaStream_t stream) {
  // dk / dv rows may be strided: dkv_row_heads heads per token row (0 = Hk, i.e. contiguous [B,S,Hk,D]).
  // With dk = base, dv = base + Hk*D and dkv_row_heads = 2*Hk both land in ONE [B,S,2*Hk*D] buffer, which lets
  // the k/v projections' dgrad and wgrad run as single GEMMs over the concatenated weight.
  if (dkv_row_heads <= 0) dkv_row_heads = Hk;
  B200_CHECK_ARG(dkv_row_heads >= Hk, "attn_bwd: dkv_row_heads=%d < Hk=%d", dkv_row_heads, Hk);
  B200_CHECK_ARG(B > 0 && S > 0 && H > 0 && Hk > 0 && H % Hk == 0,
                 "attn_bwd: bad shape B=%d S=%d H=%d Hk=%d", B, S, H, Hk);
  B200_CHECK_ARG(D == 64 || D == 128, "attn_bwd: head_dim %d unsupported (64 or 128)", D);
  if (ws_bytes < attn_bwd_workspace_bytes(B, S, H, Hk, D)) {
    set_error("attn_bwd: workspace too small (%zu < %zu)", ws_bytes,
              attn_bwd_workspace_bytes(B, S, H, Hk, D));
    return B200_ERR_WORKSPACE;
  }
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "attn_bwd: workspace must be 256-byte aligned");
  float* dq_acc = reinterpret_cast<float*>(ws);
  float* delta = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + al256((size_t)B * S * H * D * 4));
  const long long n = (long long)B * S * H * D;
  {
    const long long threads = (long long)B * S * H * (D / 8);
    const int blocks = (int)((threads + 255) / 256);
    if (D == 64)
      attn_delta_kernel<64><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o, delta, dq_acc, B, S, H);
    else
      attn_delta_kernel<128><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o, delta, dq_acc, B, S, H);
    B200_CHECK_LAUNCH();
  }
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int rc;
  if ((rc = make_bshd_map(&tmQ, q, B, S, H, D, BT, false))) return rc;
  if ((rc = make_bshd_map(&tmK, k, B, S, Hk, D, BT, false))) return rc;
  if ((rc = make_bshd_map(&tmV, v, B, S, Hk, D, BT, false))) return rc;
  if ((rc = make_bshd_map(&tmDO, d_o, B, S, H, D, BT, false))) return rc;
  BwdArgs a;
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
    // debug: B200_ATTN_TRACE=<device pointer, hex> makes CTA 0 record clock64 stamps (tools/attn_trace.py)
    const char* e = getenv("B200_ATTN_TRACE");
    a.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 16)) : nullptr;
  }
  static const bool force_v1 = [] {
    const char* e = getenv("B200_ATTN_BWD_V1");
    return e != nullptr && e[0] == '1';
  }();
  const bool pipelined = !force_v1;   // software-pipelined kernels (attention_bwd64.cu / attention_bwd128.cu)
  if (pipelined && D == 64)
    rc = launch_attn_bwd64(tmQ, tmK, tmV, tmDO, lse, delta, dq_acc, dk, dv, B, S, H, Hk, dkv_row_heads, scale, causal,
                           stream);
  else if (pipelined)
    rc = launch_attn_bwd128(tmQ, tmK, tmV, tmDO, lse, delta, dq_acc, dk, dv, B, S, H, Hk, dkv_row_heads, scale, causal,
                            stream);
  else
    rc = D == 64 ? launch_bwd<64>(tmQ, tmK, tmV, tmDO, a, stream) : launch_bwd<128>(tmQ, tmK, tmV, tmDO, a, stream);
  if (rc) return rc;
  {
    const long long nvec = n / 8;
    long long blocks = (nvec + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    // the pipelined kernels accumulate dS without the softmax scale (one multiply less per score): applied here
    f32_to_bf16_kernel<<<(int)blocks, 256, 0, stream>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(dq), nvec,
                                                        pipelined ? scale : 1.0f);
    B200_CHECK_LAUNCH();
  }
  return B200_OK;
}

}  // namespace b200
