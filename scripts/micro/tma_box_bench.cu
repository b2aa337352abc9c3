// How the L2 -> SM rate of TMA tile loads depends on the BOX a single instruction carries and on the number of issuing
// threads, for the access pattern of the UNet's K = 320 GEMMs (A [40960, 320] streamed, B [320, 320] re-read by every
// CTA; everything L2-resident).  l2_tma_bench.cu showed 16 KB boxes saturate at ~35 B/clk/SM and 32 KB boxes at ~57; this
// benchmark asks whether 3-D boxes (several 64-wide k-blocks of the same 128 rows in ONE instruction, landing as
// consecutive 128B-swizzled tiles - the layout tcgen05.mma reads) get the large-box rate.
//   mode 0: per k-block one 2-D A box (64 x 128 = 16 KB) + one 2-D B box (64 x 160 = 20 KB)     [the shipped GEMM]
//   mode 1: per 2 k-blocks one 3-D A box (64 x 128 x 2 = 32 KB) + one 3-D B box (64 x 160 x 2 = 40 KB)
//   mode 2: A only, 2-D 16 KB boxes                                                               [B-stationary, small box]
//   mode 3: A only, 3-D 32 KB boxes (2 k-blocks)
//   mode 4: A only, 3-D 80 KB boxes (the whole 128 x 320 row tile in one instruction)
//   mode 5: A only, 3-D 48 KB boxes (3 k-blocks; the last one of a K = 320 row is half out of bounds -> zero fill)
// issuers = 1 | 2 : boxes issued by one thread, or alternately by one thread each of two warps (own ring halves).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tma_box_bench.out scripts/micro/tma_box_bench.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct Args {
  int mode, stages, issuers, passes, m_tiles, kb_per_box, with_b;
};

constexpr int KB = 5;  // 64-wide k-blocks of a K = 320 row

__global__ void __launch_bounds__(128, 1)
bench_kernel(const __grid_constant__ CUtensorMap a2, const __grid_constant__ CUtensorMap b2,
             const __grid_constant__ CUtensorMap a3, const __grid_constant__ CUtensorMap b3, Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int a_bytes = a.kb_per_box * 16384, b_bytes = a.with_b ? a.kb_per_box * 20480 : 0;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane != 0 || warp >= a.issuers) return;
  // steps of this CTA: (pass, m_tile = blockIdx.x + i * gridDim.x, box index within the row tile)
  const int boxes_per_tile = (KB + a.kb_per_box - 1) / a.kb_per_box;
  const int my_tiles = (a.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int iters = a.passes * my_tiles * boxes_per_tile;
  for (int it = warp; it < iters + a.stages; it += a.issuers) {
    const int s = it % a.stages;
    if (it >= a.stages) mbar_wait(&bars[s], ((it / a.stages) - 1) & 1);
    if (it < iters) {
      const int bi = it % boxes_per_tile;
      const int mt = (int)blockIdx.x + ((it / boxes_per_tile) % my_tiles) * (int)gridDim.x;
      const int nt = (it / boxes_per_tile) & 1;
      uint8_t* sa = smem + (size_t)s * stage_bytes;
      mbar_expect_tx(&bars[s], stage_bytes);
      if (a.kb_per_box == 1) {
        tma_load_2d(sa, &a2, &bars[s], bi * 64, mt * 128);
        if (a.with_b) tma_load_2d(sa + a_bytes, &b2, &bars[s], bi * 64, nt * 160);
      } else {
        tma_load_3d(sa, &a3, &bars[s], 0, mt * 128, bi * a.kb_per_box);
        if (a.with_b) tma_load_3d(sa + a_bytes, &b3, &bars[s], 0, nt * 160, bi * a.kb_per_box);
      }
    }
  }
}

// Does a warp that issues TMA loads slow down the OTHER warps of its scheduler (SM sub-partition)?  Warp 0 (lane 0) streams
// 16 KB boxes through a 4-stage ring like bench_kernel; warps 4..7 (one per sub-partition: warp w runs on sub-partition
// w % 4) each run the same dependent-FMA loop and report their elapsed clocks.  `do_tma` = 0: warp 0 idles (baseline).
__global__ void __launch_bounds__(256, 1)
interfere_kernel(const __grid_constant__ CUtensorMap a2, int do_tma, int iters, int fma_iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stages = 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * 16384);
  __shared__ volatile int stop;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0 && do_tma) {
      int it = 0;
      for (; it < iters && !stop; ++it) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(&bars[s], ((it / stages) - 1) & 1);
        mbar_expect_tx(&bars[s], 16384);
        tma_load_2d(smem + (size_t)s * 16384, &a2, &bars[s], (it % 5) * 64, ((int)blockIdx.x + (it / 5) % 2 * 148) * 128);
      }
      // drain
      for (int k = (it > stages ? it - stages : 0); k < it; ++k) mbar_wait(&bars[k % stages], (k / stages) & 1);
      if (blockIdx.x == 0) out[8] = it;
    }
  } else if (warp >= 4) {
    float x = 1.0f + lane * 1e-3f, y = 0.5f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < fma_iters; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) x = fmaf(x, 0.999f, y);  // dependent chain: 1 warp alone issues every ~4 clk
    }
    const long long t1 = clock64();
    if (x == 123.456f) out[9] = 1;
    if (lane == 0 && blockIdx.x == 0) out[warp - 4] = t1 - t0;
    __syncwarp();
    if (warp == 4 && lane == 0) stop = 1;
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}
static CUtensorMap map2d(void* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t str[1] = {K * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode 2d failed %d\n", (int)r);
    exit(1);
  }
  return m;
}
// [rows][K] seen as (64 cols, rows, K / 64 k-blocks): a box of kb k-blocks lands as kb consecutive [box_rows][64] tiles
static CUtensorMap map3d(void* base, uint64_t rows, uint64_t K, uint32_t box_rows, uint32_t kb) {
  CUtensorMap m;
  cuuint64_t dims[3] = {64, rows, K / 64};
  cuuint64_t str[2] = {K * 2, 128};
  cuuint32_t box[3] = {64, box_rows, kb};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode 3d failed %d (box_rows %u kb %u)\n", (int)r, box_rows, kb);
    exit(1);
  }
  return m;
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int M = 40960, K = 320, N = 320;
  uint8_t *abuf, *bbuf;
  cudaMalloc(&abuf, (size_t)M * K * 2);
  cudaMalloc(&bbuf, (size_t)N * K * 2);
  cudaMemset(abuf, 1, (size_t)M * K * 2);
  cudaMemset(bbuf, 1, (size_t)N * K * 2);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("sms %d, nominal %d MHz; A [%d, %d] 16-bit (%.1f MB), B [%d, %d]\n", sms, clk / 1000, M, K, M * K * 2 / 1e6, N, K);
  {
    long long* dout;
    cudaMalloc(&dout, 16 * sizeof(long long));
    cudaFuncSetAttribute(interfere_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    CUtensorMap a2 = map2d(abuf, M, K, 128);
    for (int do_tma = 0; do_tma < 2; ++do_tma) {
      cudaMemset(dout, 0, 16 * sizeof(long long));
      interfere_kernel<<<sms, 256, 4 * 16384 + 256>>>(a2, do_tma, 1 << 20, 4000, dout);
      long long h[16];
      cudaMemcpy(h, dout, sizeof h, cudaMemcpyDeviceToHost);
      printf("interference do_tma=%d: FMA-loop clocks of the warps on sub-partitions 0..3 = %lld %lld %lld %lld (issuer on 0; boxes issued %lld) %s\n",
             do_tma, h[0], h[1], h[2], h[3], h[8], cudaGetLastError() == cudaSuccess ? "" : "ERR");
    }
  }
  struct Cfg { int mode, kb, with_b; const char* name; };
  const Cfg cfgs[] = {{0, 1, 1, "A 16K + B 20K 2-D boxes"}, {1, 2, 1, "A 32K + B 40K 3-D boxes"},
                      {2, 1, 0, "A only 16K 2-D boxes"},    {3, 2, 0, "A only 32K 3-D boxes"},
                      {5, 3, 0, "A only 48K 3-D boxes"},    {4, 5, 0, "A only 80K 3-D boxes"}};
  for (const Cfg& c : cfgs)
    for (int issuers = 1; issuers <= 2; ++issuers)
      for (int stages = 2; stages <= 6; stages += 2) {
        const int stage_bytes = c.kb * 16384 + (c.with_b ? c.kb * 20480 : 0);
        if ((size_t)stages * stage_bytes > 220 * 1024) continue;
        Args a;
        a.mode = c.mode;
        a.stages = stages;
        a.issuers = issuers;
        a.passes = 8;
        a.m_tiles = M / 128;
        a.kb_per_box = c.kb;
        a.with_b = c.with_b;
        CUtensorMap a2 = map2d(abuf, M, K, 128), b2 = map2d(bbuf, N, K, 160);
        CUtensorMap a3 = map3d(abuf, M, K, 128, c.kb), b3 = map3d(bbuf, N, K, 160, c.kb);
        const size_t smem = (size_t)stages * stage_bytes + 256;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEventRecord(e0);
          bench_kernel<<<sms, 128, smem>>>(a2, b2, a3, b3, a);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms_ = 0;
          cudaEventElapsedTime(&ms_, e0, e1);
          if (rep > 0 && ms_ < best) best = ms_;
        }
        cudaError_t err = cudaGetLastError();
        // useful bytes: the in-bounds part of every box
        const double bytes = (double)a.passes * a.m_tiles * KB * (16384 + (c.with_b ? 20480 : 0));
        printf("%-26s issuers %d stages %d: %8.3f ms  %6.2f TB/s  (%5.1f B/clk/SM at nominal) %s\n", c.name, issuers, stages,
               best, bytes / best / 1e9, bytes / sms / (best * 1e-3 * clk * 1e3), err == cudaSuccess ? "" : cudaGetErrorString(err));
      }
  return 0;
}
