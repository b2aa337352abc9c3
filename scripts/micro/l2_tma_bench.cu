// L2 -> SM (and HBM -> SM) bandwidth through TMA tile loads, and SM -> L2 through TMA stores, with nothing else running:
// the ceiling the small-K GEMMs of the UNet run against.  One persistent CTA per SM, one thread drives a ring of
// `stages` 16 KB / 32 KB boxes with mbarriers (the GEMM's producer pattern, no MMA, no epilogue).
//   mode 0: every CTA streams its own slice of a `footprint` MB buffer (footprint < 126 MB => L2 hits after warm-up)
//   mode 1: every CTA re-reads the SAME 200 KB region (the weight re-read pattern of a K = 320 GEMM)
//   mode 2: half of the boxes from the private slice, half from the shared region (A + B of a GEMM tile)
//   mode 3: TMA stores of the ring to the private slice (epilogue pattern)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/l2_tma_bench.out scripts/micro/l2_tma_bench.cu
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}

struct Args {
  int mode, stages, box_rows, iters, rows_per_cta, shared_rows;
};

// buffer viewed as [rows][64] 16-bit elements (128 B per row); a box = box_rows x 64 elements
__global__ void __launch_bounds__(128, 1)
bench_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map_shared, Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int box_bytes = a.box_rows * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * box_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int boxes_per_slice = a.rows_per_cta / a.box_rows;
  const int shared_boxes = a.shared_rows / a.box_rows;
  if (a.mode == 3) {
    for (int it = 0; it < a.iters; ++it) {
      const int s = it % a.stages;
      if (it >= a.stages) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(7) : "memory");
      tma_store_2d(&map, smem + (size_t)s * box_bytes, 0, row0 + (it % boxes_per_slice) * a.box_rows);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    return;
  }
  // prologue: fill the ring; steady state: wait oldest, re-issue
  for (int it = 0; it < a.iters + a.stages; ++it) {
    const int s = it % a.stages;
    if (it >= a.stages) mbar_wait(&bars[s], ((it / a.stages) - 1) & 1);
    if (it < a.iters) {
      mbar_expect_tx(&bars[s], box_bytes);
      const bool shared = a.mode == 1 || (a.mode == 2 && (it & 1));
      if (shared) tma_load_2d(smem + (size_t)s * box_bytes, &map_shared, &bars[s], 0, (it % shared_boxes) * a.box_rows);
      else tma_load_2d(smem + (size_t)s * box_bytes, &map, &bars[s], 0, row0 + (it % boxes_per_slice) * a.box_rows);
    }
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}
static CUtensorMap make_map(void* base, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t str[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const size_t big = (size_t)4 << 30;
  uint8_t* buf;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 1, big);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int footprints_mb[] = {32, 64, 2048};
  printf("sms %d, nominal %d MHz\n", sms, clk / 1000);
  for (int mode = 0; mode < 4; ++mode)
    for (int fi = 0; fi < 3; ++fi)
      for (int box_rows = 128; box_rows <= 256; box_rows *= 2)
        for (int stages = 2; stages <= 12; stages += (stages < 4 ? 2 : 4)) {
          if ((size_t)stages * box_rows * 128 > 200 * 1024) continue;
          if (mode == 1 && fi > 0) continue;
          Args a;
          a.mode = mode;
          a.stages = stages;
          a.box_rows = box_rows;
          const size_t fp = (size_t)footprints_mb[fi] << 20;
          a.rows_per_cta = (int)(fp / 128 / sms / 256 * 256);
          a.shared_rows = 1536 + 256 - (1536 % 256);  // ~200 KB
          const size_t total_bytes = fp * (fi == 2 ? 1 : 8);  // several passes over an L2-resident footprint
          a.iters = (int)(total_bytes / sms / ((size_t)box_rows * 128));
          CUtensorMap m = make_map(buf, (uint64_t)a.rows_per_cta * sms, box_rows);
          CUtensorMap ms = make_map(buf + ((size_t)3 << 30), a.shared_rows, box_rows);
          const size_t smem = (size_t)stages * box_rows * 128 + 256;
          float best = 1e30f;
          for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            bench_kernel<<<sms, 128, smem>>>(m, ms, a);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms_ = 0;
            cudaEventElapsedTime(&ms_, e0, e1);
            if (rep > 0 && ms_ < best) best = ms_;
          }
          cudaError_t err = cudaGetLastError();
          const double bytes = (double)a.iters * box_rows * 128 * sms;
          printf("mode %d footprint %4d MB box %2d KB stages %2d: %8.3f ms  %7.2f TB/s  (%5.1f B/clk/SM at nominal) %s\n", mode,
                 footprints_mb[fi], box_rows * 128 / 1024, stages, best, bytes / best / 1e9,
                 bytes / sms / (best * 1e-3 * clk * 1e3), err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
  return 0;
}
