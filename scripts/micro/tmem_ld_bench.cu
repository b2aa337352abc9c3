// tcgen05.ld (TMEM -> registers) throughput per SM for the shapes an epilogue can use, at 4 / 8 / 16 warps per CTA.
// Background: three epilogue-heavy kernels of this library (K = 320 GEMM, GEGLU GEMM, d = 40 flash attention) all
// settle at ~15-16 bytes of accumulator per clock per SM, whatever else is changed - this measures the ceiling.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tmem_ld_bench.out scripts/micro/tmem_ld_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_ASM(SHAPE, NUM, NREG, ...)                                                                          \
  asm volatile("tcgen05.ld.sync.aligned." SHAPE "." NUM ".b32 {" __VA_ARGS__ "}, [%" #NREG "];" : REGS##NREG : "r"(taddr) : "memory")

template <int MODE> __device__ __forceinline__ uint32_t do_ld(uint32_t taddr) {
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = 0;
  if constexpr (MODE == 0) {  // 32x32b.x16 (what the GEMM epilogue uses): 16 regs, 2 KB per warp instruction
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
  } else if constexpr (MODE == 1) {  // 32x32b.x32: 32 regs, 4 KB
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
  } else if constexpr (MODE == 2) {  // 32x32b.x8: 8 regs, 1 KB
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
  } else if constexpr (MODE == 3) {  // 16x256b.x4: 16 regs, 16 lanes x 32 columns = 2 KB
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
  } else if constexpr (MODE == 4) {  // 16x128b.x8: 16 regs, 16 lanes x 32 columns = 2 KB
    asm volatile(
        "tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
  } else if constexpr (MODE == 5) {  // 16x64b.x16: 16 regs, 16 lanes x 32 columns = 2 KB
    asm volatile(
        "tcgen05.ld.sync.aligned.16x64b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= r[i];
  return s;
}

template <int MODE> __global__ void bench(uint32_t* out, long long* cycles, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) acc ^= do_ld<MODE>(base + (uint32_t)((it * 32 + (warp >> 2) * 64) & 255));
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE> void run(const char* name, int bytes_per_warp_instr, uint32_t* out, long long* cyc) {
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 4096;
    bench<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    bench<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148.0;
    const double bytes = (double)iters * warps * bytes_per_warp_instr;
    printf("%-14s warps %2d: %8.0f clk for %6.1f KB per SM -> %6.1f B/clk/SM, %6.1f clk per warp instruction  %s\n", name,
           warps, avg, bytes / 1024.0, bytes / avg, avg / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<2>("32x32b.x8", 1024, out, cyc);
  run<0>("32x32b.x16", 2048, out, cyc);
  run<1>("32x32b.x32", 4096, out, cyc);
  run<3>("16x256b.x4", 2048, out, cyc);
  run<4>("16x128b.x8", 2048, out, cyc);
  run<5>("16x64b.x16", 2048, out, cyc);
  return 0;
}
