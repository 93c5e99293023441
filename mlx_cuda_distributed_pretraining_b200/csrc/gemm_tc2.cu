// 2-CTA (cta_group::2) variant of the batched bf16 GEMM engine in gemm_tc.cu.
//
// Why: with 128 x 256 single-CTA tiles every CTA pulls 48 KB per k-block through L2 for 4.2 MFLOP
// (87 flop/B).  At ~6.3 KB/cycle of L2->SM bandwidth for the whole chip that caps the kernel near
// 1.0 PFLOP/s (ncu r01: tensor pipe 60 % active, DRAM 12 %).  A CTA pair on one TPC computing a
// 256 x 256 tile with tcgen05.mma.cta_group::2 loads A[128 x 64] + B[128 x 64] per CTA (B halves are
// read by both tensor cores), i.e. 64 KB per 8.4 MFLOP = 131 flop/B.
//
// Pair protocol (leader = cluster rank 0):
//   both CTAs   TMA-load their A rows and their half of B (cp.async.bulk.tensor ... cta_group::2,
//               completion bytes land on the LEADER's full barrier), epilogue their 128 rows
//   leader      waits full[stage] (2 producer arrivals + both CTAs' bytes), issues the MMAs for the
//               pair, tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs
//   tmem_empty  lives in the leader; the peer's epilogue warps arrive remotely (mapa + arrive)
#include "common.cuh"
#include "host.h"

namespace b200 {

namespace {

constexpr int G2_EPI_WARPS = 8;  // two warps per TMEM lane quarter, each owning half of the tile's columns
constexpr int G2_THREADS = 128 + 32 * G2_EPI_WARPS;
constexpr int G2_BK = 64;
constexpr int G2_A_BYTES = 128 * G2_BK * 2;  // 16 KB per CTA per stage

struct Gemm2Args {
  int M, N, K, batch;
  int tiles_m, tiles_n;  // pair tiles: 256 x BN
  float alpha, beta;
  const float* alpha_vec;
  const float* beta_vec;
  const void* C;
  long long ldc, strideC;
  void* D;
  long long ldd, strideD;
  int symmetric;  // D is symmetric (M == N, BN == 256): compute tiles with mi <= ni, mirror-write the rest
  int k_splits;   // > 1: every tile's K range is cut into k_splits work items, each storing its raw fp32
  float* ws;      //      partial into slab ws[split][batch][M][N]; splitk_finalize_kernel applies the epilogue
  // GEMM -> all-gather in one kernel: besides D, the epilogue stores every output tile into the same offsets
  // of `n_peers` peer-mapped copies of D (other GPUs' buffers over NVLink; torch symmetric memory).  Tiles
  // leave through a per-warp smem transpose so that every store instruction writes 8 rows x 64 contiguous
  // bytes (full sectors / NVLink flits) instead of 32 rows x 16 bytes.  bf16 output, non-symmetric only.
  int n_peers;
  void* peer_D[7];
  int staged_epi;  // use the coalescing (smem-transposed) epilogue even without peers
  // fused MLP epilogues (EPI template parameter; arch/llama.py:149-151 of the reference):
  //   EPI 1  [g | u] = x [Wg ; Wu]^T with y = g * sigmoid(u) * 2 computed from the fp32 accumulators: pair tile =
  //          128 gate columns (leader CTA's B rows) + the SAME 128 up columns (peer CTA's B rows at +glu_I);
  //          D = gu [M, 2*glu_I] (ldd = 2*glu_I), aux = y [M, glu_I]
  //   EPI 2  d = dy Wd (dgrad of down_proj) with the GLU adjoint applied to the accumulator: aux = gu (read),
  //          D = [dg | du] [M, 2*glu_I], dg = d*sigmoid(u)*2, du = d*g*sigmoid(u)*(1-sigmoid(u))*2
  int glu_I;
  void* aux;
  long long ld_aux;
};
constexpr int G2_MAX_PEERS = 7;

// tile index -> (mi, ni).  Symmetric mode walks the upper triangle row by row.
__device__ __forceinline__ void tile_coords(int symmetric, int tiles_n, int r, int& mi, int& ni) {
  if (!symmetric) {
    mi = r / tiles_n;
    ni = r - mi * tiles_n;
    return;
  }
  int row = 0, len = tiles_n;
  while (r >= len) {
    r -= len;
    ++row;
    --len;
  }
  mi = row;
  ni = row + r;
}

template <int BN>
struct G2Cfg {
  static constexpr int BNH = BN / 2;  // B rows held by each CTA
  static constexpr int B_BYTES = BNH * G2_BK * 2;
  static constexpr int STAGE_BYTES = G2_A_BYTES + B_BYTES;        // per CTA
  static constexpr int STAGES = (BN == 256) ? 6 : 8;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int EPI_STAGE_BYTES = 8 * 4096;  // 8 epilogue warps x ([32 rows x 64 B] + its transpose for mirror tiles)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_STAGE_BYTES;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `target` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t target) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(target)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0,
                                             int c1, int c2) {
  // peer bit cleared: completion bytes are credited to the leader CTA's barrier
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}


// Everything the epilogue needs for one work item (one 256 x BN pair tile, or one K-split of it).
struct G2Tile {
  int M, N, batch;
  int b, mi, ni, split, k_splits;
  float alpha, beta;  // per-batch vectors already folded in
  const void* C;
  long long ldc, strideC;
  void* D;
  long long ldd, strideD;
  float* ws;
  bool mirror;  // symmetric-output mode, off-diagonal tile: also write the transposed tile
  bool staged;  // coalescing (smem-transposed) stores; required for peer stores
  int n_peers;
  void* const* peer_D;
};

// One epilogue warp's share of one work item: TMEM accumulator -> alpha/beta/C -> global (and peers).
template <typename OutT, int BN>
__device__ __forceinline__ void g2_epilogue_tile(const G2Tile& t, uint32_t tmem_base, uint32_t acc, uint32_t acc_phase,
                                                 uint32_t tfull, uint32_t tempty, uint32_t epi_stage, int warp, int lane,
                                                 uint32_t rank) {
  constexpr int VEC = (sizeof(OutT) == 2) ? 8 : 4;
  constexpr int BNW = BN / (G2_EPI_WARPS / 4);  // columns per epilogue warp
  constexpr int NCV = BNW / VEC;                // 16-byte C vectors per row and warp
  const int q = warp & 3;
  const int cbase = ((warp - 4) >> 2) * BNW;
  const int row = q * 32 + lane;
  const int b = t.b;
  const int m0 = t.mi * 256 + (int)rank * 128;
  const int n0 = t.ni * BN;
  const bool mirror = t.mirror;
  if (t.k_splits > 1) {
    // split-K: raw fp32 partial into this split's slab; alpha/beta/C/mirroring happen in the finalize
    const int gm_ = m0 + row;
    float* wrow = t.ws + (((long long)t.split * t.batch + b) * t.M + gm_) * (long long)t.N;
    mbar_wait(tfull, acc_phase);
    tc_fence_after_sync();
    const uint32_t t_row_ = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 2
    for (int c0 = cbase; c0 < cbase + BNW; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row_ + c0, v);
      tmem_ld_wait();
      if (gm_ < t.M) {
#pragma unroll
        for (int g = 0; g < 32; g += 4) {
          const int gn = n0 + c0 + g;
          if (gn < t.N)
            *reinterpret_cast<float4*>(wrow + gn) =
                make_float4(__uint_as_float(v[g]), __uint_as_float(v[g + 1]), __uint_as_float(v[g + 2]),
                            __uint_as_float(v[g + 3]));
        }
      }
    }
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(tempty, 0);
    return;
  }
  const float alpha = t.alpha;
  const float beta = t.beta;
  const int gm = m0 + row;
  const bool row_ok = gm < t.M;
  OutT* drow = reinterpret_cast<OutT*>(t.D) + (long long)b * t.strideD + (long long)gm * t.ldd;
  const OutT* crow =
      t.C ? reinterpret_cast<const OutT*>(t.C) + (long long)b * t.strideC + (long long)gm * t.ldc : nullptr;

  // bf16 C: fetch the whole row of the epilogue input BEFORE the accumulator is ready, so its
  // latency hides behind the MMAs of this tile instead of serialising the epilogue
  uint4 cpre[(sizeof(OutT) == 2) ? NCV : 1];
  if constexpr (sizeof(OutT) == 2) {
    if (crow && row_ok) {
#pragma unroll
      for (int i = 0; i < NCV; ++i)
        cpre[i] = (n0 + cbase + i * 8 < t.N) ? ldg128(crow + n0 + cbase + i * 8) : make_uint4(0, 0, 0, 0);
    }
  }

  mbar_wait(tfull, acc_phase);
  tc_fence_after_sync();
  const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll
  for (int cc = 0; cc < BNW; cc += 32) {
    const int c0 = cbase + cc;
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + c0, v);
    tmem_ld_wait();
    if constexpr (VEC == 8) {
      if (t.staged) {
        // ---- fused all-gather path: alpha/beta epilogue -> smem transpose -> coalesced local + peer stores
        const uint32_t stg = epi_stage + (uint32_t)(warp - 4) * 4096u;
        const uint32_t stgT = stg + 2048u;  // [32 columns][32 rows]: the sub-tile transposed (mirror tiles)
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = alpha * __uint_as_float(v[g + i]);
          if (crow && row_ok) {
            const uint4 cv = cpre[(cc + g) / 8];
            const float2 c01 = unpack_bf16x2(cv.x), c23 = unpack_bf16x2(cv.y), c45 = unpack_bf16x2(cv.z),
                         c67 = unpack_bf16x2(cv.w);
            f[0] = fmaf(beta, c01.x, f[0]);
            f[1] = fmaf(beta, c01.y, f[1]);
            f[2] = fmaf(beta, c23.x, f[2]);
            f[3] = fmaf(beta, c23.y, f[3]);
            f[4] = fmaf(beta, c45.x, f[4]);
            f[5] = fmaf(beta, c45.y, f[5]);
            f[6] = fmaf(beta, c67.x, f[6]);
            f[7] = fmaf(beta, c67.y, f[7]);
          }
          const int piece = g >> 3;
          const uint32_t addr = stg + lane * 64 + (((piece ^ ((lane >> 1) & 3))) << 4);
          const uint32_t w4[4] = {pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                  pack_bf16x2(f[6], f[7])};
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]),
                       "r"(w4[3])
                       : "memory");
          if (mirror) {
            // transposed copy: element (row = lane, col j) -> [j][lane]; the 32 lanes of one store fill one
            // 64-byte row of the transposed tile (conflict-free), 16-byte pieces swizzled like the direct tile
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = g + i;
              const uint16_t h = (i & 1) ? (uint16_t)(w4[i >> 1] >> 16) : (uint16_t)(w4[i >> 1] & 0xffffu);
              const uint32_t ta = stgT + j * 64 + ((((lane >> 3) ^ ((j >> 1) & 3))) << 4) + (lane & 7) * 2;
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(ta), "h"(h) : "memory");
            }
          }
        }
        __syncwarp();
        const int piece = lane & 3;
        const int gn = n0 + c0 + piece * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2);  // row within this warp's 32
          uint4 o;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                       : "r"(stg + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4)));
          const int grow = m0 + q * 32 + r;
          if (grow < t.M && gn < t.N) {
            const long long off = (long long)b * t.strideD + (long long)grow * t.ldd + gn;
            stg128(reinterpret_cast<OutT*>(t.D) + off, o);
            for (int pr = 0; pr < t.n_peers; ++pr) stg128(reinterpret_cast<OutT*>(t.peer_D[pr]) + off, o);
          }
        }
        if (mirror) {
          // D[n0 + c0 + j][m0 + 32 q + 8 piece ..] = transposed sub-tile rows: 8 rows x 64 B per instruction
          // (the first version issued 32 two-byte stores per lane here)
          const int mcol = m0 + q * 32 + piece * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = i * 8 + (lane >> 2);
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                         : "r"(stgT + j * 64 + ((piece ^ ((j >> 1) & 3)) << 4)));
            const int mrow = n0 + c0 + j;
            if (mrow < t.M && mcol < t.N)
              stg128(reinterpret_cast<OutT*>(t.D) + (long long)b * t.strideD + (long long)mrow * t.ldd + mcol, o);
          }
        }
        __syncwarp();
        continue;
      }
    }
    if (row_ok) {
#pragma unroll
      for (int g = 0; g < 32; g += VEC) {
        const int gn = n0 + c0 + g;
        if (gn < t.N) {
          float f[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) f[i] = alpha * __uint_as_float(v[g + i]);
          if constexpr (VEC == 8) {
            if (crow) {
              const uint4 cv = cpre[(cc + g) / 8];
              const float2 c01 = unpack_bf16x2(cv.x), c23 = unpack_bf16x2(cv.y), c45 = unpack_bf16x2(cv.z),
                           c67 = unpack_bf16x2(cv.w);
              f[0] = fmaf(beta, c01.x, f[0]);
              f[1] = fmaf(beta, c01.y, f[1]);
              f[2] = fmaf(beta, c23.x, f[2]);
              f[3] = fmaf(beta, c23.y, f[3]);
              f[4] = fmaf(beta, c45.x, f[4]);
              f[5] = fmaf(beta, c45.y, f[5]);
              f[6] = fmaf(beta, c67.x, f[6]);
              f[7] = fmaf(beta, c67.y, f[7]);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            stg128(drow + gn, o);
            if (mirror) {
              // D[gn + i][gm] = D[gm][gn + i]: for a fixed column the 32 lanes of the warp hold 32
              // consecutive rows, so each 2-byte store instruction fills one contiguous 64-byte run
              __nv_bfloat16* dcol = reinterpret_cast<__nv_bfloat16*>(t.D) + (long long)b * t.strideD +
                                    (long long)gn * t.ldd + gm;
              const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint16_t h = (i & 1) ? (uint16_t)(w[i >> 1] >> 16) : (uint16_t)(w[i >> 1] & 0xffffu);
                *reinterpret_cast<uint16_t*>(dcol + (long long)i * t.ldd) = h;
              }
            }
          } else {
            if (crow) {
              const float4 cv = *reinterpret_cast<const float4*>(crow + gn);
              f[0] = fmaf(beta, cv.x, f[0]);
              f[1] = fmaf(beta, cv.y, f[1]);
              f[2] = fmaf(beta, cv.z, f[2]);
              f[3] = fmaf(beta, cv.w, f[3]);
            }
            *reinterpret_cast<float4*>(drow + gn) = make_float4(f[0], f[1], f[2], f[3]);
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(tempty, 0);
}


// sigmoid on the special-function unit with ONE MUFU op: sigmoid(x) = 0.5 + 0.5 tanh(x / 2) (tanh.approx.f32, abs error
// ~5e-4 of a value in [0, 1] that is rounded to bf16 anyway).  The GLU adjoint's epilogue evaluates 32 768 sigmoids
// per CTA tile: with ex2 + rcp (two MUFU ops, 16 per clock per SM) that alone matched the tile's 4096 MMA cycles.
__device__ __forceinline__ float g2_sigmoid(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// Per-warp staging tile [32 rows x 64 B] (16-byte pieces XOR-swizzled by row pair, conflict-free both ways): lanes own
// ROWS when they produce / consume values (TMEM lane = row) but a global access instruction should cover few rows and
// whole 64-byte runs -- 8 rows x 64 B per instruction instead of 32 rows x 16 B (4x fewer LSU wavefronts / L2 requests).
__device__ __forceinline__ void g2_stage_put_row(uint32_t stg, int lane, const uint32_t* w16) {
#pragma unroll
  for (int piece = 0; piece < 4; ++piece)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4)),
                 "r"(w16[piece * 4 + 0]), "r"(w16[piece * 4 + 1]), "r"(w16[piece * 4 + 2]), "r"(w16[piece * 4 + 3])
                 : "memory");
}
__device__ __forceinline__ void g2_stage_get_row(uint32_t stg, int lane, uint32_t* w16) {
#pragma unroll
  for (int piece = 0; piece < 4; ++piece)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w16[piece * 4 + 0]), "=r"(w16[piece * 4 + 1]), "=r"(w16[piece * 4 + 2]), "=r"(w16[piece * 4 + 3])
                 : "r"(stg + lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4)));
}
// staged tile -> global: rows row0 .. row0+31 of a row-major bf16 matrix, 32 columns starting at col0
__device__ __forceinline__ void g2_stage_store(uint32_t stg, int lane, __nv_bfloat16* base, long long ld, int row0,
                                               int col0, int M, int ncols) {
  const int piece = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    uint4 o;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                 : "r"(stg + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4)));
    if (row0 + r < M && col0 + piece * 8 < ncols) stg128(base + (long long)(row0 + r) * ld + col0 + piece * 8, o);
  }
}
// global -> registers in the coalesced pattern (4 x 16 B per lane); g2_stage_fill parks them in the staging tile
__device__ __forceinline__ void g2_coalesced_load(const __nv_bfloat16* base, long long ld, int row0, int col0, int M,
                                                  int ncols, int lane, uint4* o4) {
  const int piece = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    o4[i] = (row0 + r < M && col0 + piece * 8 < ncols) ? ldg128(base + (long long)(row0 + r) * ld + col0 + piece * 8)
                                                       : make_uint4(0, 0, 0, 0);
  }
}
__device__ __forceinline__ void g2_stage_fill(uint32_t stg, int lane, const uint4* o4) {
  const int piece = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4)),
                 "r"(o4[i].x), "r"(o4[i].y), "r"(o4[i].z), "r"(o4[i].w)
                 : "memory");
  }
}

// EPI 1: one epilogue warp = 32 rows x 64 gate columns + the matching 64 up columns of a [256 x (128 | 128)] pair tile
__device__ __forceinline__ void g2_epilogue_glu_fwd(const Gemm2Args& p, int b, int mi, int ni, uint32_t tmem_base,
                                                    uint32_t acc, uint32_t acc_phase, uint32_t tfull, uint32_t tempty,
                                                    uint32_t epi_stage, int warp, int lane, uint32_t rank) {
  constexpr int BN = 256;
  const int q = warp & 3;
  const int hsel = (warp - 4) >> 2;  // which 64-column half of the 128 gate / up columns
  const int row0 = mi * 256 + (int)rank * 128 + q * 32;
  const int col0 = ni * 128 + hsel * 64;  // feature index of this warp's first column
  __nv_bfloat16* gu = reinterpret_cast<__nv_bfloat16*>(p.D) + (long long)b * p.strideD;
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.aux);
  const uint32_t stg = epi_stage + (uint32_t)(warp - 4) * 4096u;  // two 2 KB tiles per warp
  mbar_wait(tfull, acc_phase);
  tc_fence_after_sync();
  const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll
  for (int cc = 0; cc < 64; cc += 32) {
    uint32_t vg[32], vu[32];
    tmem_ld_32x32b_x32(t_row + hsel * 64 + cc, vg);
    tmem_ld_32x32b_x32(t_row + 128 + hsel * 64 + cc, vu);
    tmem_ld_wait();
    uint32_t wg[16], wu[16], wy[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float g0 = __uint_as_float(vg[2 * i]), g1 = __uint_as_float(vg[2 * i + 1]);
      const float u0 = __uint_as_float(vu[2 * i]), u1 = __uint_as_float(vu[2 * i + 1]);
      wg[i] = pack_bf16x2(g0, g1);
      wu[i] = pack_bf16x2(u0, u1);
      wy[i] = pack_bf16x2(g0 * g2_sigmoid(u0) * 2.0f, g1 * g2_sigmoid(u1) * 2.0f);
    }
    g2_stage_put_row(stg, lane, wg);
    g2_stage_put_row(stg + 2048, lane, wu);
    __syncwarp();
    g2_stage_store(stg, lane, gu, p.ldd, row0, col0 + cc, p.M, p.glu_I);
    g2_stage_store(stg + 2048, lane, gu + p.glu_I, p.ldd, row0, col0 + cc, p.M, p.glu_I);
    __syncwarp();
    g2_stage_put_row(stg, lane, wy);
    __syncwarp();
    g2_stage_store(stg, lane, yb, p.ld_aux, row0, col0 + cc, p.M, p.glu_I);
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(tempty, 0);
}

// EPI 2: standard [256 x 256] pair tile of d = dy Wd; one epilogue warp = 32 rows x 128 columns
__device__ __forceinline__ void g2_epilogue_glu_bwd(const Gemm2Args& p, int b, int mi, int ni, uint32_t tmem_base,
                                                    uint32_t acc, uint32_t acc_phase, uint32_t tfull, uint32_t tempty,
                                                    uint32_t epi_stage, int warp, int lane, uint32_t rank) {
  constexpr int BN = 256;
  const int q = warp & 3;
  const int cbase = ((warp - 4) >> 2) * 128;
  const int row0 = mi * 256 + (int)rank * 128 + q * 32;
  const int n0 = ni * BN + cbase;
  const __nv_bfloat16* gu = reinterpret_cast<const __nv_bfloat16*>(p.aux);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.D) + (long long)b * p.strideD;
  const uint32_t stg = epi_stage + (uint32_t)(warp - 4) * 4096u;
  // the first chunk's g / u are requested (coalesced: 8 rows x 64 B per instruction) before the accumulator is ready
  uint4 pg[4], pu[4];
  g2_coalesced_load(gu, p.ld_aux, row0, n0, p.M, p.glu_I, lane, pg);
  g2_coalesced_load(gu + p.glu_I, p.ld_aux, row0, n0, p.M, p.glu_I, lane, pu);
  mbar_wait(tfull, acc_phase);
  tc_fence_after_sync();
  const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 1
  for (int cc = 0; cc < 128; cc += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + cbase + cc, v);
    g2_stage_fill(stg, lane, pg);
    g2_stage_fill(stg + 2048, lane, pu);
    if (cc + 32 < 128) {  // next chunk's g / u in flight while this one is processed
      g2_coalesced_load(gu, p.ld_aux, row0, n0 + cc + 32, p.M, p.glu_I, lane, pg);
      g2_coalesced_load(gu + p.glu_I, p.ld_aux, row0, n0 + cc + 32, p.M, p.glu_I, lane, pu);
    }
    __syncwarp();
    uint32_t gw[16], uw[16];
    g2_stage_get_row(stg, lane, gw);
    g2_stage_get_row(stg + 2048, lane, uw);
    tmem_ld_wait();
    uint32_t w1[16], w2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 gg = unpack_bf16x2(gw[i]), uu = unpack_bf16x2(uw[i]);
      const float d0 = __uint_as_float(v[2 * i]), d1 = __uint_as_float(v[2 * i + 1]);
      const float s0 = g2_sigmoid(uu.x), s1 = g2_sigmoid(uu.y);
      w1[i] = pack_bf16x2(d0 * s0 * 2.0f, d1 * s1 * 2.0f);
      w2[i] = pack_bf16x2(d0 * gg.x * s0 * (1.0f - s0) * 2.0f, d1 * gg.y * s1 * (1.0f - s1) * 2.0f);
    }
    __syncwarp();
    g2_stage_put_row(stg, lane, w1);
    g2_stage_put_row(stg + 2048, lane, w2);
    __syncwarp();
    g2_stage_store(stg, lane, out, p.ldd, row0, n0 + cc, p.M, p.glu_I);
    g2_stage_store(stg + 2048, lane, out + p.glu_I, p.ldd, row0, n0 + cc, p.M, p.glu_I);
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(tempty, 0);
}

template <bool A_MN, bool B_MN, typename OutT, int BN, int EPI = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const Gemm2Args p) {
  using Cfg = G2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BNH = Cfg::BNH;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t epi_stage = bar_base + 256;  // after the barrier block
  auto smem_a = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + G2_A_BYTES; };

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  cluster_sync_all();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);   // one arrival per CTA's producer (used in the leader only)
      mbar_init(empty_bar(s), 1);  // multicast commit from the leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);   // multicast commit
      mbar_init(tempty_bar(a), 2 * G2_EPI_WARPS);  // epilogue warps x 2 CTAs (leader only)
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem2_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();

  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch)
  // may overlap the tail of the previous kernel in the stream; nothing below may, because operands and
  // C/D belong to the chain.  Our own dependents are released right away - they block in the same wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int tiles_per_batch = p.symmetric ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
  const int total_tiles = tiles_per_batch * p.batch;
  const int num_kb = (p.K + G2_BK - 1) / G2_BK;
  const int total_work = total_tiles * p.k_splits;  // work item w -> (tile w / k_splits, split w % k_splits)

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------- TMA producer (both CTAs) -----------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int tile = w / p.k_splits;
        const int split = w - tile * p.k_splits;
        const int kb_begin = (int)((long long)split * num_kb / p.k_splits);
        const int kb_end = (int)((long long)(split + 1) * num_kb / p.k_splits);
        const int b = tile / tiles_per_batch;
        int mi, ni;
        tile_coords(p.symmetric, p.tiles_n, tile - b * tiles_per_batch, mi, ni);
        const int m0 = mi * 256 + (int)rank * 128;
        // EPI 1: the leader loads 128 rows of Wg, the peer the SAME 128 rows of Wu (stacked glu_I rows further down)
        const int n0 = (EPI == 1) ? ni * 128 + (int)rank * p.glu_I : ni * BN + (int)rank * BNH;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const int k0 = kb * G2_BK;
          if constexpr (!A_MN) {
            tma2_load_3d(smem_a(stage), &tmA, full_bar(stage), k0, m0, b);
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i)
              tma2_load_3d(smem_a(stage) + i * 8192, &tmA, full_bar(stage), m0 + i * 64, k0, b);
          }
          if constexpr (!B_MN) {
            tma2_load_3d(smem_b(stage), &tmB, full_bar(stage), k0, n0, b);
          } else {
#pragma unroll
            for (int i = 0; i < BNH / 64; ++i)
              tma2_load_3d(smem_b(stage) + i * 8192, &tmB, full_bar(stage), n0 + i * 64, k0, b);
          }
          if (leader)
            mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          else
            mbar_arrive_cluster(full_bar(stage), 0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ------------- MMA issuer (leader only): the whole warp runs the (uniform) control flow, -----
      // ------------- one elected lane issues, so descriptors stay on the uniform datapath ---------
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t iter = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters, ++iter) {
        const int split = w % p.k_splits;
        const int kb_begin = (int)((long long)split * num_kb / p.k_splits);
        const int kb_end = (int)((long long)(split + 1) * num_kb / p.k_splits);
        const uint32_t acc = iter & 1u;
        const uint32_t acc_phase = (iter >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_a(stage);
          const uint32_t b_addr = smem_b(stage);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k) {
              const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                       : make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
              const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                       : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
              umma2_bf16_ss(d_tmem, da, db, idesc, (kb != kb_begin || k != 0) ? 1u : 0u);
            }
            umma2_commit_mc(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma2_commit_mc(tfull_bar(acc));
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue (both CTAs) -----------------------------------
    uint32_t iter = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters, ++iter) {
      const int tile = w / p.k_splits;
      const int b = tile / tiles_per_batch;
      if constexpr (EPI != 0) {
        int mi, ni;
        tile_coords(0, p.tiles_n, tile - b * tiles_per_batch, mi, ni);
        if constexpr (EPI == 1)
          g2_epilogue_glu_fwd(p, b, mi, ni, tmem_base, iter & 1u, (iter >> 1) & 1u, tfull_bar(iter & 1u),
                              tempty_bar(iter & 1u), epi_stage, warp, lane, rank);
        else
          g2_epilogue_glu_bwd(p, b, mi, ni, tmem_base, iter & 1u, (iter >> 1) & 1u, tfull_bar(iter & 1u),
                              tempty_bar(iter & 1u), epi_stage, warp, lane, rank);
        continue;
      }
      G2Tile t;
      t.M = p.M;
      t.N = p.N;
      t.batch = p.batch;
      t.b = b;
      tile_coords(p.symmetric, p.tiles_n, tile - b * tiles_per_batch, t.mi, t.ni);
      t.split = w - tile * p.k_splits;
      t.k_splits = p.k_splits;
      t.alpha = p.alpha * (p.alpha_vec ? p.alpha_vec[b] : 1.0f);
      t.beta = p.beta * (p.beta_vec ? p.beta_vec[b] : 1.0f);
      t.C = p.C;
      t.ldc = p.ldc;
      t.strideC = p.strideC;
      t.D = p.D;
      t.ldd = p.ldd;
      t.strideD = p.strideD;
      t.ws = p.ws;
      t.mirror = p.symmetric && t.mi != t.ni;
      t.staged = p.n_peers > 0 || p.staged_epi;
      t.n_peers = p.n_peers;
      t.peer_D = p.peer_D;
      g2_epilogue_tile<OutT, BN>(t, tmem_base, iter & 1u, (iter >> 1) & 1u, tfull_bar(iter & 1u), tempty_bar(iter & 1u),
                                 epi_stage, warp, lane, rank);
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem2_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// Split-K epilogue: D = bf16(alpha * sum_s ws[s]) in a fixed summation order (deterministic), with the
// symmetric mode's mirror writes.  Lanes own consecutive rows, as in the GEMM epilogue, so the mirrored
// 2-byte stores of one instruction fill a contiguous 64-byte run.
__global__ void __launch_bounds__(256)
splitk_finalize_kernel(const float* __restrict__ ws, int splits, int batch, int M, int N, __nv_bfloat16* D,
                       long long ldd, long long strideD, float alpha, const float* __restrict__ alpha_vec,
                       int symmetric) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const int j0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * 8;
  if (i >= M || j0 >= N) return;
  const int ti = i >> 8, tj = j0 >> 8;
  if (symmetric && ti > tj) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = 0; s < splits; ++s) {
    const float* src = ws + (((long long)s * batch + b) * M + i) * (long long)N + j0;
    const float4 v0 = *reinterpret_cast<const float4*>(src);
    const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
    acc[0] += v0.x; acc[1] += v0.y; acc[2] += v0.z; acc[3] += v0.w;
    acc[4] += v1.x; acc[5] += v1.y; acc[6] += v1.z; acc[7] += v1.w;
  }
  const float a = alpha * (alpha_vec ? alpha_vec[b] : 1.0f);
  uint4 o;
  o.x = pack_bf16x2(a * acc[0], a * acc[1]);
  o.y = pack_bf16x2(a * acc[2], a * acc[3]);
  o.z = pack_bf16x2(a * acc[4], a * acc[5]);
  o.w = pack_bf16x2(a * acc[6], a * acc[7]);
  __nv_bfloat16* dbase = D + (long long)b * strideD;
  stg128(dbase + (long long)i * ldd + j0, o);
  if (symmetric && ti != tj) {
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const uint16_t h = (t & 1) ? (uint16_t)(w[t >> 1] >> 16) : (uint16_t)(w[t >> 1] & 0xffffu);
      *reinterpret_cast<uint16_t*>(dbase + (long long)(j0 + t) * ldd + i) = h;
    }
  }
}


// ------------------------------------------------------------------------------------------------------
// Grouped launch: up to G2G_MAX independent batched GEMM problems (different M/N/K/batch/layouts/epilogues)
// served by ONE persistent grid.  Used by the Newton-Schulz chain: one launch per stage (A = X X^T, B = bA + cAA,
// X' = aX + BX) across ALL shape groups of the model instead of one per group, because a single group's 240
// symmetric tiles fill 3.24 waves of the 74 CTA pairs (19 % of the launch idles in the tail) and the chain pays
// launch + pipeline-fill latency 75 times per optimizer step.  bf16 output, 256-wide pair tiles; operand layouts
// are per-problem run-time flags (descriptor major bits + TMA box shapes), everything else as in the kernel above.
// Work items (tile, K-split) of all problems are concatenated in the order given by the host (heaviest K first)
// and dealt round-robin to the clusters; a problem whose tiles are much heavier than the rest (the lone
// 32003 x 1024 embedding's X^T X) is cut along K into items of about the common size.
// ------------------------------------------------------------------------------------------------------
constexpr int G2G_MAX = 6;

struct G2GProblem {
  int M, N, K, batch;
  int tiles_n, tiles_per_batch;
  int num_kb;
  int work_begin;  // first work item of this problem in the concatenated list
  int a_mn, b_mn, symmetric, k_splits, staged_epi, n_peers;
  float alpha, beta;
  const float* alpha_vec;
  const float* beta_vec;
  const void* C;
  long long ldc, strideC;
  void* D;
  long long ldd, strideD;
  float* ws;
  void* peer_D[G2_MAX_PEERS];
};

struct G2GArgs {
  int n_problems;
  int total_work;
  G2GProblem p[G2G_MAX];
  CUtensorMap tmA[G2G_MAX];
  CUtensorMap tmB[G2G_MAX];
};

__device__ __forceinline__ int g2g_find(const G2GArgs& g, int w) {
  int pi = 0;
#pragma unroll
  for (int i = 1; i < G2G_MAX; ++i)
    if (i < g.n_problems && w >= g.p[i].work_begin) pi = i;
  return pi;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_grouped_kernel(const __grid_constant__ G2GArgs g) {
  constexpr int BN = 256;
  using Cfg = G2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BNH = Cfg::BNH;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t epi_stage = bar_base + 256;
  auto smem_a = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + G2_A_BYTES; };

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  cluster_sync_all();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < g.n_problems; ++i) {
      tma_prefetch_desc(&g.tmA[i]);
      tma_prefetch_desc(&g.tmB[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * G2_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem2_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();

  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int total_work = g.total_work;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------- TMA producer (both CTAs) -----------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int pi = g2g_find(g, w);
        const G2GProblem& q = g.p[pi];
        const CUtensorMap* mA = &g.tmA[pi];
        const CUtensorMap* mB = &g.tmB[pi];
        const int local = w - q.work_begin;
        const int tile = local / q.k_splits;
        const int split = local - tile * q.k_splits;
        const int kb_begin = (int)((long long)split * q.num_kb / q.k_splits);
        const int kb_end = (int)((long long)(split + 1) * q.num_kb / q.k_splits);
        const int b = tile / q.tiles_per_batch;
        int mi, ni;
        tile_coords(q.symmetric, q.tiles_n, tile - b * q.tiles_per_batch, mi, ni);
        const int m0 = mi * 256 + (int)rank * 128;
        const int n0 = ni * BN + (int)rank * BNH;
        const bool a_mn = q.a_mn != 0, b_mn = q.b_mn != 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const int k0 = kb * G2_BK;
          if (!a_mn) {
            tma2_load_3d(smem_a(stage), mA, full_bar(stage), k0, m0, b);
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) tma2_load_3d(smem_a(stage) + i * 8192, mA, full_bar(stage), m0 + i * 64, k0, b);
          }
          if (!b_mn) {
            tma2_load_3d(smem_b(stage), mB, full_bar(stage), k0, n0, b);
          } else {
#pragma unroll
            for (int i = 0; i < BNH / 64; ++i)
              tma2_load_3d(smem_b(stage) + i * 8192, mB, full_bar(stage), n0 + i * 64, k0, b);
          }
          if (leader)
            mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          else
            mbar_arrive_cluster(full_bar(stage), 0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ------------- MMA issuer (leader only) -----------------------------------------------------
      int stage = 0;
      uint32_t phase = 0;
      uint32_t iter = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters, ++iter) {
        const int pi = g2g_find(g, w);
        const G2GProblem& q = g.p[pi];
        const int local = w - q.work_begin;
        const int split = local % q.k_splits;
        const int kb_begin = (int)((long long)split * q.num_kb / q.k_splits);
        const int kb_end = (int)((long long)(split + 1) * q.num_kb / q.k_splits);
        const bool a_mn = q.a_mn != 0, b_mn = q.b_mn != 0;
        const uint32_t idesc = make_idesc_bf16(256, BN, a_mn, b_mn);
        // per-k-step advance of the operand descriptors: 32 B inside the swizzle row (K-major) or two 1 KB
        // swizzle atoms (MN-major); leading-dimension byte offset 0 / 8192 respectively
        const uint32_t a_step = a_mn ? 2048u : 32u, a_lbo = a_mn ? 8192u : 0u;
        const uint32_t b_step = b_mn ? 2048u : 32u, b_lbo = b_mn ? 8192u : 0u;
        const uint32_t acc = iter & 1u;
        const uint32_t acc_phase = (iter >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_a(stage);
          const uint32_t b_addr = smem_b(stage);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k) {
              const uint64_t da = make_smem_desc_sw128(a_addr + k * a_step, a_lbo, 1024);
              const uint64_t db = make_smem_desc_sw128(b_addr + k * b_step, b_lbo, 1024);
              umma2_bf16_ss(d_tmem, da, db, idesc, (kb != kb_begin || k != 0) ? 1u : 0u);
            }
            umma2_commit_mc(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma2_commit_mc(tfull_bar(acc));
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue (both CTAs) -----------------------------------
    uint32_t iter = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters, ++iter) {
      const int pi = g2g_find(g, w);
      const G2GProblem& q = g.p[pi];
      const int local = w - q.work_begin;
      const int tile = local / q.k_splits;
      const int b = tile / q.tiles_per_batch;
      G2Tile t;
      t.M = q.M;
      t.N = q.N;
      t.batch = q.batch;
      t.b = b;
      tile_coords(q.symmetric, q.tiles_n, tile - b * q.tiles_per_batch, t.mi, t.ni);
      t.split = local - tile * q.k_splits;
      t.k_splits = q.k_splits;
      t.alpha = q.alpha * (q.alpha_vec ? q.alpha_vec[b] : 1.0f);
      t.beta = q.beta * (q.beta_vec ? q.beta_vec[b] : 1.0f);
      t.C = q.C;
      t.ldc = q.ldc;
      t.strideC = q.strideC;
      t.D = q.D;
      t.ldd = q.ldd;
      t.strideD = q.strideD;
      t.ws = q.ws;
      t.mirror = q.symmetric && t.mi != t.ni;
      t.staged = q.n_peers > 0 || q.staged_epi;
      t.n_peers = q.n_peers;
      t.peer_D = q.peer_D;
      g2_epilogue_tile<__nv_bfloat16, BN>(t, tmem_base, iter & 1u, (iter >> 1) & 1u, tfull_bar(iter & 1u),
                                          tempty_bar(iter & 1u), epi_stage, warp, lane, rank);
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem2_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN, typename OutT, int BN, int EPI = 0>
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const Gemm2Args& args, cudaStream_t stream) {
  using Cfg = G2Cfg<BN>;
  auto kern = gemm2_bf16_tc_kernel<A_MN, B_MN, OutT, BN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = (args.symmetric ? args.tiles_m * (args.tiles_m + 1) / 2 : args.tiles_m * args.tiles_n) *
                    args.batch * args.k_splits;
  const int pairs = num_sms() / 2;
  const int clusters = total < pairs ? total : pairs;
  static const bool use_pdl = [] {
    const char* e = getenv("B200_GEMM_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  B200_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, args));
  B200_CHECK_LAUNCH();
  return B200_OK;
}

template <bool A_MN, bool B_MN>
int dispatch2(bool out_f32, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const Gemm2Args& a,
              cudaStream_t s) {
  if (out_f32)
    return bn == 256 ? launch_gemm2<A_MN, B_MN, float, 256>(tmA, tmB, a, s)
                     : launch_gemm2<A_MN, B_MN, float, 128>(tmA, tmB, a, s);
  return bn == 256 ? launch_gemm2<A_MN, B_MN, __nv_bfloat16, 256>(tmA, tmB, a, s)
                   : launch_gemm2<A_MN, B_MN, __nv_bfloat16, 128>(tmA, tmB, a, s);
}

}  // namespace


// ---- fused MLP GEMMs (arch/llama.py:142-151: down(gate(x) * sigmoid(up(x)) * 2)) --------------------------
static int glu_common_args(Gemm2Args& a, int M, int N_tiles_n, int K, int I, void* D, long long ldd, void* aux,
                           long long ld_aux) {
  a = Gemm2Args{};
  a.M = M;
  a.N = I;
  a.K = K;
  a.batch = 1;
  a.tiles_m = (M + 255) / 256;
  a.tiles_n = N_tiles_n;
  a.alpha = 1.0f;
  a.beta = 0.0f;
  a.D = D;
  a.ldd = ldd;
  a.strideD = 0;
  a.k_splits = 1;
  a.glu_I = I;
  a.aux = aux;
  a.ld_aux = ld_aux;
  return 0;
}

// gu[M, 2I] = x[M, K] [Wg ; Wu]^T (W2: [2I, K] row-major, gate rows first), y[M, I] = g * sigmoid(u) * 2
int gemm_glu_fwd_2cta(const void* x, const void* W2, void* gu, void* y, int M, int K, int I, cudaStream_t stream) {
  B200_CHECK_ARG(M > 0 && K > 0 && I > 0 && K % 8 == 0 && I % 128 == 0,
                 "mlp_glu_fwd: need K %% 8 == 0 and intermediate size %% 128 == 0 (M=%d K=%d I=%d)", M, K, I);
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, 1};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)M * K * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = make_tensor_map(&tmA, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)(2 * I), 1};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)2 * I * K * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = make_tensor_map(&tmB, W2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  Gemm2Args a;
  glu_common_args(a, M, I / 128, K, I, gu, 2LL * I, y, I);
  return launch_gemm2<false, false, __nv_bfloat16, 256, 1>(tmA, tmB, a, stream);
}

// dgu[M, 2I] = GLU adjoint of d = dy[M, H] Wd (Wd: [H, I] row-major = down_proj.weight), gu[M, 2I] the saved
// forward activations: dg = d * sigmoid(u) * 2 ; du = d * g * sigmoid(u) * (1 - sigmoid(u)) * 2
int gemm_glu_bwd_2cta(const void* dy, const void* Wd, const void* gu, void* dgu, int M, int H, int I,
                      cudaStream_t stream) {
  B200_CHECK_ARG(M > 0 && H > 0 && I > 0 && H % 8 == 0 && I % 8 == 0, "mlp_glu_bwd: bad shape M=%d H=%d I=%d", M, H, I);
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)H, (uint64_t)M, 1};
    const uint64_t strides[2] = {(uint64_t)H * 2, (uint64_t)M * H * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = make_tensor_map(&tmA, dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  {  // B[n = i, k = h] = Wd[h, i]: stored [K = H rows][N = I]: MN-major
    const uint64_t dims[3] = {(uint64_t)I, (uint64_t)H, 1};
    const uint64_t strides[2] = {(uint64_t)I * 2, (uint64_t)H * I * 2};
    const uint32_t box[3] = {64, (uint32_t)G2_BK, 1};
    int rc = make_tensor_map(&tmB, Wd, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  Gemm2Args a;
  glu_common_args(a, M, (I + 255) / 256, H, I, dgu, 2LL * I, const_cast<void*>(gu), 2LL * I);
  return launch_gemm2<false, true, __nv_bfloat16, 256, 2>(tmA, tmB, a, stream);
}

// ---- grouped launch, host side --------------------------------------------------------------------
// Problems are bf16-out, 256-wide-tile CTA-pair GEMMs (M > 128).  `k_splits` of a problem > 1 requests split-K
// into fp32 slabs (splitk_ws) + a finalize pass, exactly as gemm_bf16_2cta does for a single problem.
int gemm_grouped_2cta(const GroupedGemm* probs, int n, cudaStream_t stream) {
  B200_CHECK_ARG(n >= 1 && n <= G2G_MAX, "grouped gemm: %d problems (1..%d supported)", n, G2G_MAX);
  static const int staged_env = [] {
    const char* e = getenv("B200_GEMM_STAGED");
    return (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }();
  // order: heaviest work item first, so that the tail of the launch is made of the cheapest tiles
  int order[G2G_MAX];
  for (int i = 0; i < n; ++i) order[i] = i;
  auto item_kb = [&](int i) {
    const int num_kb = (probs[i].K + G2_BK - 1) / G2_BK;
    const int ks = probs[i].k_splits > 1 ? probs[i].k_splits : 1;
    return (num_kb + ks - 1) / ks;
  };
  for (int i = 1; i < n; ++i)
    for (int j = i; j > 0 && item_kb(order[j]) > item_kb(order[j - 1]); --j) {
      const int t = order[j];
      order[j] = order[j - 1];
      order[j - 1] = t;
    }
  G2GArgs a;
  a.n_problems = n;
  int work = 0;
  for (int oi = 0; oi < n; ++oi) {
    const GroupedGemm& h = probs[order[oi]];
    G2GProblem& q = a.p[oi];
    B200_CHECK_ARG(h.M > 128 && h.N > 0 && h.K > 0 && h.batch > 0 && h.N % 8 == 0,
                   "grouped gemm: problem %d has an unsupported shape M=%d N=%d K=%d batch=%d", order[oi], h.M, h.N,
                   h.K, h.batch);
    B200_CHECK_ARG(h.lda % 8 == 0 && h.ldb % 8 == 0 && h.ldd % 8 == 0 && (h.C == nullptr || h.ldc % 8 == 0),
                   "grouped gemm: leading dimensions must be multiples of 8");
    B200_CHECK_ARG(h.n_peers >= 0 && h.n_peers <= G2_MAX_PEERS, "grouped gemm: n_peers out of range");
    int symmetric = (h.symmetric && h.M == h.N) ? 1 : 0;
    int k_splits = h.k_splits > 1 ? h.k_splits : 1;
    if (k_splits > 1 && (h.splitk_ws == nullptr || h.beta != 0.0f)) k_splits = 1;
    B200_CHECK_ARG(h.n_peers == 0 || (!symmetric && k_splits == 1),
                   "grouped gemm: peer stores need a non-symmetric, non-split-K problem");
    {
      const uint64_t dims[3] = {(uint64_t)(h.a_mn ? h.M : h.K), (uint64_t)(h.a_mn ? h.K : h.M), (uint64_t)h.batch};
      const uint64_t strides[2] = {(uint64_t)h.lda * 2,
                                   (uint64_t)(h.batch > 1 ? h.strideA : (long long)dims[1] * h.lda) * 2};
      const uint32_t box[3] = {64, (uint32_t)(h.a_mn ? G2_BK : 128), 1};
      int rc = make_tensor_map(&a.tmA[oi], h.A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, true);
      if (rc) return rc;
    }
    {
      const uint64_t dims[3] = {(uint64_t)(h.b_mn ? h.N : h.K), (uint64_t)(h.b_mn ? h.K : h.N), (uint64_t)h.batch};
      const uint64_t strides[2] = {(uint64_t)h.ldb * 2,
                                   (uint64_t)(h.batch > 1 ? h.strideB : (long long)dims[1] * h.ldb) * 2};
      const uint32_t box[3] = {64, (uint32_t)(h.b_mn ? G2_BK : 128), 1};
      int rc = make_tensor_map(&a.tmB[oi], h.B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, true);
      if (rc) return rc;
    }
    const int tiles_m = (h.M + 255) / 256;
    q.M = h.M;
    q.N = h.N;
    q.K = h.K;
    q.batch = h.batch;
    q.tiles_n = (h.N + 255) / 256;
    q.tiles_per_batch = symmetric ? tiles_m * (tiles_m + 1) / 2 : tiles_m * q.tiles_n;
    q.num_kb = (h.K + G2_BK - 1) / G2_BK;
    q.work_begin = work;
    q.a_mn = h.a_mn ? 1 : 0;
    q.b_mn = h.b_mn ? 1 : 0;
    q.symmetric = symmetric;
    q.k_splits = k_splits;
    q.staged_epi = (staged_env != 0 && !(staged_env == 2 && symmetric) && k_splits <= 1) ? 1 : 0;
    q.n_peers = h.n_peers;
    q.alpha = h.alpha;
    q.beta = h.beta;
    q.alpha_vec = h.alpha_vec;
    q.beta_vec = h.beta_vec;
    q.C = (h.beta != 0.0f) ? h.C : nullptr;
    q.ldc = h.ldc;
    q.strideC = h.strideC;
    q.D = h.D;
    q.ldd = h.ldd;
    q.strideD = h.strideD;
    q.ws = h.splitk_ws;
    for (int i = 0; i < G2_MAX_PEERS; ++i) q.peer_D[i] = i < h.n_peers ? const_cast<void*>(h.peer_D[i]) : nullptr;
    work += q.tiles_per_batch * h.batch * k_splits;
  }
  a.total_work = work;
  auto kern = gemm2_grouped_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<256>::SMEM_BYTES));
    attr_set = true;
  }
  const int pairs = num_sms() / 2;
  const int clusters = work < pairs ? work : pairs;
  static const bool use_pdl = [] {
    const char* e = getenv("B200_GEMM_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = G2Cfg<256>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  B200_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  B200_CHECK_LAUNCH();
  for (int oi = 0; oi < n; ++oi) {
    const G2GProblem& q = a.p[oi];
    if (q.k_splits <= 1) continue;
    const dim3 grid((q.M + 31) / 32, (q.N + 63) / 64, q.batch);
    splitk_finalize_kernel<<<grid, 256, 0, stream>>>(q.ws, q.k_splits, q.batch, q.M, q.N,
                                                     reinterpret_cast<__nv_bfloat16*>(q.D), q.ldd, q.strideD, q.alpha,
                                                     q.alpha_vec, q.symmetric);
    B200_CHECK_LAUNCH();
  }
  return B200_OK;
}

// Same contract as gemm_bf16 (gemm_tc.cu); arguments are assumed validated by the caller.
int gemm_bf16_2cta(bool a_mn, bool b_mn, int M, int N, int K, int batch, const void* A, long long lda,
                   long long strideA, const void* B, long long ldb, long long strideB, const void* C,
                   long long ldc, long long strideC, void* D, long long ldd, long long strideD, bool out_f32,
                   float alpha, float beta, const float* alpha_vec, const float* beta_vec, int bn,
                   int symmetric, int k_splits, float* splitk_ws, const void* const* peer_D, int n_peers,
                   cudaStream_t stream) {
  if (symmetric && (M != N || bn != 256 || out_f32)) symmetric = 0;  // square bf16 256-tiles only
  B200_CHECK_ARG(n_peers >= 0 && n_peers <= G2_MAX_PEERS, "gemm: n_peers=%d out of range (0..%d)", n_peers,
                 G2_MAX_PEERS);
  B200_CHECK_ARG(n_peers == 0 || (!out_f32 && !symmetric && k_splits <= 1),
                 "gemm: peer stores need a bf16, non-symmetric, non-split-K output");
  // split-K needs a plain alpha-only bf16 epilogue and 16-byte partial rows
  if (k_splits > 1 && (!splitk_ws || out_f32 || beta != 0.0f || (N & 7) != 0)) k_splits = 1;
  if (k_splits < 1) k_splits = 1;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)(a_mn ? M : K), (uint64_t)(a_mn ? K : M), (uint64_t)batch};
    const uint64_t strides[2] = {(uint64_t)lda * 2, (uint64_t)(batch > 1 ? strideA : (long long)dims[1] * lda) * 2};
    const uint32_t box[3] = {64, (uint32_t)(a_mn ? G2_BK : 128), 1};
    int rc = make_tensor_map(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)(b_mn ? N : K), (uint64_t)(b_mn ? K : N), (uint64_t)batch};
    const uint64_t strides[2] = {(uint64_t)ldb * 2, (uint64_t)(batch > 1 ? strideB : (long long)dims[1] * ldb) * 2};
    const uint32_t box[3] = {64, (uint32_t)(b_mn ? G2_BK : bn / 2), 1};
    int rc = make_tensor_map(&tmB, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  Gemm2Args a;
  a.M = M;
  a.N = N;
  a.K = K;
  a.batch = batch;
  a.tiles_m = (M + 255) / 256;
  a.tiles_n = (N + bn - 1) / bn;
  a.alpha = alpha;
  a.beta = beta;
  a.alpha_vec = alpha_vec;
  a.beta_vec = beta_vec;
  a.C = (beta != 0.0f) ? C : nullptr;
  a.ldc = ldc;
  a.strideC = strideC;
  a.D = D;
  a.ldd = ldd;
  a.strideD = strideD;
  a.symmetric = symmetric;
  a.k_splits = k_splits;
  a.ws = splitk_ws;
  // the coalescing epilogue is also the faster one for plain local stores (8 rows x 64 B per instruction
  // instead of 32 rows x 16 B: Newton-Schulz chain -3 %); B200_GEMM_STAGED=0 restores direct stores
  // 0 = direct stores, 1 = staged everywhere, 2 = staged except symmetric tiles (default: measured A/B on one
  // box, Newton-Schulz chain 3.62 ms vs 3.66 ms with the transposed-mirror variant, 3.64 -> 3.53 vs mode 0)
  static const int staged_env = [] {
    const char* e = getenv("B200_GEMM_STAGED");
    return (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }();
  a.staged_epi = (staged_env != 0 && !(staged_env == 2 && symmetric) && !out_f32 && k_splits <= 1) ? 1 : 0;
  a.n_peers = n_peers;
  for (int i = 0; i < G2_MAX_PEERS; ++i) a.peer_D[i] = i < n_peers ? const_cast<void*>(peer_D[i]) : nullptr;
  a.glu_I = 0;
  a.aux = nullptr;
  a.ld_aux = 0;
  int rc;
  if (!a_mn && !b_mn) rc = dispatch2<false, false>(out_f32, bn, tmA, tmB, a, stream);
  else if (!a_mn && b_mn) rc = dispatch2<false, true>(out_f32, bn, tmA, tmB, a, stream);
  else if (a_mn && b_mn) rc = dispatch2<true, true>(out_f32, bn, tmA, tmB, a, stream);
  else rc = dispatch2<true, false>(out_f32, bn, tmA, tmB, a, stream);
  if (rc || k_splits == 1) return rc;
  const dim3 grid((M + 31) / 32, (N + 63) / 64, batch);
  splitk_finalize_kernel<<<grid, 256, 0, stream>>>(splitk_ws, k_splits, batch, M, N,
                                                   reinterpret_cast<__nv_bfloat16*>(D), ldd, strideD, alpha,
                                                   alpha_vec, symmetric);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace b200
