// SM -> L2 store bandwidth with ordinary vector stores, to set beside the 22-24 B/clk/SM that TMA tile stores reach
// (l2_tma_bench.cu mode 3): is the GEMM epilogue's store path (staging tile -> TMA store) limited by the TMA engine or by
// the SM's store port?  One persistent CTA per SM, `warps` warps, every lane writes 16 bytes per instruction.
//   pattern 0: 512 contiguous bytes per warp instruction (upper bound)
//   pattern 1: row segments of 320 bytes at a 640-byte pitch (the 128 x 160 output tile of an N = 320 GEMM): 20 lanes
//              per row, a warp instruction covers 1.6 rows
//   pattern 2: row segments of 320 bytes at a 1920-byte pitch (N = 960)
// footprint: 32 MB (stays in L2, rewritten 8 times) or 2048 MB (streams to HBM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/store_bench.out scripts/micro/store_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(512, 1)
store_kernel(uint4* base, size_t bytes_per_cta, int passes, int pattern, int pitch) {
  uint4* my = base + (size_t)blockIdx.x * (bytes_per_cta / 16);
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3u, 4u);
  const size_t n16 = bytes_per_cta / 16;
  for (int p = 0; p < passes; ++p) {
    if (pattern == 0) {
      for (size_t i = threadIdx.x; i < n16; i += blockDim.x) my[i] = v;
    } else {
      // logical index i -> (row, 16-byte column within the 320-byte segment)
      const size_t rows = bytes_per_cta / pitch;
      for (size_t i = threadIdx.x; i < rows * 20; i += blockDim.x) {
        const size_t r = i / 20, c = i - r * 20;
        my[r * (pitch / 16) + c] = v;
      }
    }
  }
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const size_t big = (size_t)3 << 30;
  uint4* buf;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 0, big);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("sms %d, nominal %d MHz\n", sms, clk / 1000);
  const int fps[] = {32, 2048};
  const int pitches[] = {512, 640, 1920};
  for (int pattern = 0; pattern < 3; ++pattern)
    for (int fi = 0; fi < 2; ++fi)
      for (int warps = 4; warps <= 16; warps *= 2) {
        const int pitch = pitches[pattern];
        size_t per_cta = ((size_t)fps[fi] << 20) / sms / 7680 * 7680;  // multiple of every pitch and of 512
        const int passes = fi == 0 ? 8 : 1;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEventRecord(e0);
          store_kernel<<<sms, warps * 32>>>(buf, per_cta, passes, pattern, pitch);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          if (rep > 0 && ms < best) best = ms;
        }
        const double written = pattern == 0 ? (double)per_cta * passes * sms
                                            : (double)(per_cta / pitch) * 320.0 * passes * sms;
        printf("pattern %d pitch %4d footprint %4d MB warps %2d: %8.3f ms  %6.2f TB/s  (%5.1f B/clk/SM at nominal) %s\n", pattern,
               pitch, fps[fi], warps, best, written / best / 1e9, written / sms / (best * 1e-3 * clk * 1e3),
               cudaGetLastError() == cudaSuccess ? "" : "ERR");
      }
  return 0;
}
