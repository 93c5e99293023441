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
};
constexpr int G2_MAX_PEERS = 7;

// tile index -> (mi, ni).  Symmetric mode walks the upper triangle row by row.
__device__ __forceinline__ void tile_coords(const Gemm2Args& p, int r, int& mi, int& ni) {
  if (!p.symmetric) {
    mi = r / p.tiles_n;
    ni = r - mi * p.tiles_n;
    return;
  }
  int row = 0, len = p.tiles_n;
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

template <bool A_MN, bool B_MN, typename OutT, int BN>
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
        tile_coords(p, tile - b * tiles_per_batch, mi, ni);
        const int m0 = mi * 256 + (int)rank * 128;
        const int n0 = ni * BN + (int)rank * BNH;
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
    constexpr int VEC = (sizeof(OutT) == 2) ? 8 : 4;
    constexpr int BNW = BN / (G2_EPI_WARPS / 4);  // columns per epilogue warp
    constexpr int NCV = BNW / VEC;                // 16-byte C vectors per row and warp
    const int q = warp & 3;
    const int cbase = ((warp - 4) >> 2) * BNW;
    const int row = q * 32 + lane;
    uint32_t iter = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters, ++iter) {
      const int tile = w / p.k_splits;
      const int split = w - tile * p.k_splits;
      const int b = tile / tiles_per_batch;
      int mi, ni;
      tile_coords(p, tile - b * tiles_per_batch, mi, ni);
      const int m0 = mi * 256 + (int)rank * 128;
      const int n0 = ni * BN;
      const bool mirror = p.symmetric && mi != ni;
      const uint32_t acc = iter & 1u;
      const uint32_t acc_phase = (iter >> 1) & 1u;
      if (p.k_splits > 1) {
        // split-K: raw fp32 partial into this split's slab; alpha/beta/C/mirroring happen in the finalize
        const int gm_ = m0 + row;
        float* wrow = p.ws + (((long long)split * p.batch + b) * p.M + gm_) * (long long)p.N;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after_sync();
        const uint32_t t_row_ = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 2
        for (int c0 = cbase; c0 < cbase + BNW; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row_ + c0, v);
          tmem_ld_wait();
          if (gm_ < p.M) {
#pragma unroll
            for (int g = 0; g < 32; g += 4) {
              const int gn = n0 + c0 + g;
              if (gn < p.N)
                *reinterpret_cast<float4*>(wrow + gn) =
                    make_float4(__uint_as_float(v[g]), __uint_as_float(v[g + 1]), __uint_as_float(v[g + 2]),
                                __uint_as_float(v[g + 3]));
            }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
        continue;
      }
      const float alpha = p.alpha * (p.alpha_vec ? p.alpha_vec[b] : 1.0f);
      const float beta = p.beta * (p.beta_vec ? p.beta_vec[b] : 1.0f);
      const int gm = m0 + row;
      const bool row_ok = gm < p.M;
      OutT* drow = reinterpret_cast<OutT*>(p.D) + (long long)b * p.strideD + (long long)gm * p.ldd;
      const OutT* crow =
          p.C ? reinterpret_cast<const OutT*>(p.C) + (long long)b * p.strideC + (long long)gm * p.ldc : nullptr;

      // bf16 C: fetch the whole row of the epilogue input BEFORE the accumulator is ready, so its
      // latency hides behind the MMAs of this tile instead of serialising the epilogue
      uint4 cpre[(sizeof(OutT) == 2) ? NCV : 1];
      if constexpr (sizeof(OutT) == 2) {
        if (crow && row_ok) {
#pragma unroll
          for (int i = 0; i < NCV; ++i)
            cpre[i] = (n0 + cbase + i * 8 < p.N) ? ldg128(crow + n0 + cbase + i * 8) : make_uint4(0, 0, 0, 0);
        }
      }

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll
      for (int cc = 0; cc < BNW; cc += 32) {
        const int c0 = cbase + cc;
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c0, v);
        tmem_ld_wait();
        if constexpr (VEC == 8) {
          if (p.n_peers > 0 || p.staged_epi) {
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
              if (grow < p.M && gn < p.N) {
                const long long off = (long long)b * p.strideD + (long long)grow * p.ldd + gn;
                stg128(reinterpret_cast<OutT*>(p.D) + off, o);
                for (int pr = 0; pr < p.n_peers; ++pr) stg128(reinterpret_cast<OutT*>(p.peer_D[pr]) + off, o);
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
                if (mrow < p.M && mcol < p.N)
                  stg128(reinterpret_cast<OutT*>(p.D) + (long long)b * p.strideD + (long long)mrow * p.ldd + mcol, o);
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
            if (gn < p.N) {
              float f[VEC];
#pragma unroll
              for (int i = 0; i < VEC; ++i) f[i] = alpha * __uint_as_float(v[g + i]);
              if constexpr (VEC == 8) {
                if (crow) {
                  const uint4 cv = cpre[(cc + g) / 8];
                  const float2 c01 = unpack_bf16x2(cv.x), c23 = unpack_bf16x2(cv.y),
                               c45 = unpack_bf16x2(cv.z), c67 = unpack_bf16x2(cv.w);
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
                  __nv_bfloat16* dcol = reinterpret_cast<__nv_bfloat16*>(p.D) + (long long)b * p.strideD +
                                        (long long)gn * p.ldd + gm;
                  const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const uint16_t h = (i & 1) ? (uint16_t)(w[i >> 1] >> 16) : (uint16_t)(w[i >> 1] & 0xffffu);
                    *reinterpret_cast<uint16_t*>(dcol + (long long)i * p.ldd) = h;
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
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
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

template <bool A_MN, bool B_MN, typename OutT, int BN>
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const Gemm2Args& args, cudaStream_t stream) {
  using Cfg = G2Cfg<BN>;
  auto kern = gemm2_bf16_tc_kernel<A_MN, B_MN, OutT, BN>;
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
