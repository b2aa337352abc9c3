// Measures MUFU.EX2 throughput (fp32 and packed f16x2) per SM per clock on the current GPU.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__global__ void k_f32(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_h2(float* out, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x2c002c00u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
  }
  unsigned s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int dev_clk;
  cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_f32<<<148, warps * 32>>>(out, iters);
        else k_h2<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double ops = 148.0 * warps * 32 * 8.0 * iters * (mode ? 2 : 1);
      printf("%s warps/SM=%2d: %.3f ms  %.2f exp/clk/SM (at %d MHz nominal)\n", mode ? "f16x2" : "f32  ", warps, ms,
             ops / 148.0 / (ms * 1e-3 * dev_clk * 1e3), dev_clk / 1000);
    }
  return 0;
}
