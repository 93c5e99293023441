// Micro-benchmark: per-SM and chip-wide throughput of the three dQ drain access patterns of the attention
// backward kernel, into an L2-resident [rows][H*D] fp32 accumulator (row stride 4096 B, 64 floats per row used):
//   A  red.v4, one warp instruction = 2 rows x 256 contiguous bytes     (smem-staged, current kernel)
//   B  red.v2, one warp instruction = 8 rows x 32 contiguous bytes      (tcgen05.ld 16x256b fragment, no staging)
//   C  red.v4, one warp instruction = 32 rows x 16 bytes                (row per lane, first version)
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ROW_FLOATS = 1024;  // H * D

__global__ void pat_kernel(float* acc, int tiles_per_cta, int n_tiles_total, int mode) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 4 warps = one drain warpgroup
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int tile = (blockIdx.x * tiles_per_cta + t) % n_tiles_total;  // 128 rows x 64 floats, head = tile % 16
    float* base = acc + (long long)(tile / 16) * 128 * ROW_FLOATS + (tile % 16) * 64;
    if (mode == 0) {
      for (int i = 0; i < 16; ++i) {
        const int c = threadIdx.x + 128 * i;  // 16-byte chunk id
        const int r = c >> 4, ch = c & 15;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(base + (long long)r * ROW_FLOATS + ch * 4),
                     "f"(1.0f) : "memory");
      }
    } else if (mode == 1) {
      for (int half = 0; half < 2; ++half)
        for (int cg = 0; cg < 8; ++cg)
          for (int hi = 0; hi < 2; ++hi) {
            const int r = warp * 32 + half * 16 + hi * 8 + lane / 4;
            const int col = cg * 8 + (lane % 4) * 2;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(base + (long long)r * ROW_FLOATS + col), "f"(1.0f)
                         : "memory");
          }
    } else {
      const int r = threadIdx.x;
      for (int ch = 0; ch < 16; ++ch)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(base + (long long)r * ROW_FLOATS + ch * 4),
                     "f"(1.0f) : "memory");
    }
  }
}

int main() {
  const int n_tiles = 16 * 8 * 16;  // B * (S/128) * H  = 2048 tiles of 32 KB = 64 MB accumulator
  float* d;
  cudaMalloc(&d, (long long)n_tiles / 16 * 128 * ROW_FLOATS * 4);
  cudaMemset(d, 0, (long long)n_tiles / 16 * 128 * ROW_FLOATS * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[] = {"A v4 2x256B", "B v2 8x32B ", "C v4 32x16B"};
  for (int mode = 0; mode < 3; ++mode)
    for (int g : {148, 16}) {
      const int tiles_per_cta = 64;
      pat_kernel<<<g, 128>>>(d, tiles_per_cta, n_tiles, mode);
      cudaEventRecord(e0);
      pat_kernel<<<g, 128>>>(d, tiles_per_cta, n_tiles, mode);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = (double)g * tiles_per_cta * 32768;
      printf("%s grid %3d: %.1f GB/s, %.1f B/clk/SM, %.0f cycles per 32 KB tile (1.9 GHz)\n", names[mode], g,
             bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1.9e9 / g, ms * 1e-3 * 1.9e9 / tiles_per_cta);
    }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
