// Micro-benchmark: throughput of coalesced red.global.add.v4.f32 vs plain st.global.v4 over a buffer
// larger than L2, for different numbers of participating CTAs.  Decides whether the attention-backward
// dQ drain is limited per SM or by the L2 atomic units chip-wide.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void red_kernel(float* dst, long long n4, int mode) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float* a = dst + i * 4;
    if (mode == 0)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(a), "f"(1.0f) : "memory");
    else if (mode == 1)
      asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(a), "f"(1.0f) : "memory");
    else
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a), "f"(1.0f) : "memory");
  }
}

int main() {
  const long long bytes = 302ll << 20;
  float* d;
  cudaMalloc(&d, bytes);
  cudaMemset(d, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grids[] = {148 * 8, 148 * 2, 148, 74, 32, 8};
  for (int mode = 0; mode < 2; ++mode)
    for (int g : grids) {
      red_kernel<<<g, 256>>>(d, bytes / 16, mode);
      cudaEventRecord(e0);
      for (int r = 0; r < 5; ++r) red_kernel<<<g, 256>>>(d, bytes / 16, mode);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%s grid %5d: %.3f ms per pass, %.1f GB/s (%.1f B/clk/CTA at 1.9 GHz)\n", mode ? "st.v4 " : "red.v4", g,
             ms / 5, bytes / (ms / 5 * 1e-3) / 1e9, bytes / (ms / 5 * 1e-3) / 1.9e9 / (g < 148 ? g : 148));
    }
  // same-address-reuse case: 4.7 MB window (L2 resident), like dq tiles revisited by several CTAs
  const long long small = 4ll << 20;
  for (int g : {148 * 8, 148}) {
    cudaEventRecord(e0);
    for (int r = 0; r < 50; ++r) red_kernel<<<g, 256>>>(d, small / 16, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("red.v4 L2-resident 4 MB grid %d: %.1f GB/s\n", g, small * 50 / (ms * 1e-3) / 1e9);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
