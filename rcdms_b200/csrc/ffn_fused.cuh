// Fused GEGLU feed-forward for the C = 320 (64x64-latent) transformer blocks (attention.py:434-436, 523-526;
// motion_module.py:231-232, 244-246):
//
//     y <- y + ( GEGLU( LayerNorm(y) W1^T + b1 ) ) W2^T + b2
//
// as ONE kernel.  The [M, 4C] intermediate (105 MB per layer at 64x64 latents, written and read back by the two separate
// GEMMs at the SM -> L2 store rate of 6.4 TB/s) never leaves the SM: a CTA owns a 128-row tile of y, keeps it resident in
// shared memory (A1, 80 KB), and walks the 4C = 1280 inner columns in chunks of 32:
//
//   GEMM1  acc1[128 x 64]  = A1[128 x 320] * W1f[chunk: 32 h rows | 32 gate rows]^T     (tcgen05, TMEM, double-buffered)
//   EPI1   h[128 x 32]     = (rstd * acc1_h + c_h) * gelu(rstd * acc1_g + c_g)  -> 16-bit -> shared memory (A2)
//   GEMM2  acc2[128 x 320] += A2[128 x 32] * W2[:, chunk]^T                               (TMEM, lives across the 40 chunks)
//   EPI2   y = round(acc2 + b2) + y                                                      (after the last chunk)
//
// LayerNorm is folded around GEMM1 exactly as in gemm_tcgen05.cuh (centred, gamma-scaled weights W1f; per-row rstd from the
// producer's row statistics; constant vector c).  W1f and c are packed with GEGLU tile width 64 (h | gate per chunk).
// Operand layouts: A1 / W1f chunks are 128B-swizzled K-major tiles written by TMA (W1f: one 3-D box = all five k-blocks of a
// chunk, 40 KB, in ONE instruction, two stages; a ring of ten single k-block boxes was measured 25 % slower: ~300 cycles of
// issue per box); W2 chunks are 64B-swizzled K-major tiles (32 columns = 64 bytes per row); A2 is written
// by the epilogue threads in the no-swizzle interleaved layout [8-column chunk][row][16 B] (conflict-free 16-byte stores).
// Warp roles (12 warps): 0 = TMA producer A1 + W1f, 1 = GEMM1 issuer, 2-9 = epilogue (TMEM lane quarter x column half),
// 10 = TMA producer W2, 11 = GEMM2 issuer.  TMEM: acc2 columns [0, 320), three acc1 buffers at 320 / 384 / 448 (all 512 columns).
// Three acc1 / A2 buffers: the epilogue of a chunk (TMEM load, GEGLU, 16-bit store, proxy fence: ~1 200-2 000 cycles for the
// slowest warp) may lag two chunks behind GEMM1 before anything waits for it.
#pragma once
#include "common.cuh"

// Timeline instrumentation (variant builds only, -DRCDM_FFN_TRACE=1; scripts/ffn_trace.py)
#ifndef RCDM_FFN_TRACE
#define RCDM_FFN_TRACE 0
#endif

namespace rcdm {

#if RCDM_FFN_TRACE
constexpr int FFN_TRACE_CHUNKS = 128;
__device__ long long g_ffn_trace[FFN_TRACE_CHUNKS * 16];
#define FFN_STAMP(g, slot)                                                                  \
  do {                                                                                      \
    if (blockIdx.x == 7 && (g) < FFN_TRACE_CHUNKS) g_ffn_trace[(g) * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define FFN_STAMP(g, slot) do { } while (0)
#endif

struct FfnParams {
  int M, num_m_tiles;
  const float2* stats_in;  // [stats_parts][M] (sum, sum of squares) of the rows of y
  int stats_parts;
  float ln_eps;
  const float* c1;     // [2J] folded-LayerNorm constant vector, packed per chunk: [32 h | 32 gate]
  const float* bias2;  // [C]
  const void* res;     // [M, C] residual (the input y; may alias out)
  void* out;           // [M, C]
};
struct FfnMaps {
  CUtensorMap a1;  // y   [M, C]      box (64, 128)        SWIZZLE_128B
  CUtensorMap b1;  // W1f [2J, C] as (64, 2J, C/64)  box (64, 64, C/64)  SWIZZLE_128B
  CUtensorMap b2;  // W2  [C, J]       box (32, 160)        SWIZZLE_64B
};

struct FfnCfg {
  static constexpr int C = 320, J = 4 * C, KB1 = C / 64;   // 5 k-blocks of GEMM1
  static constexpr int CH = 32;                            // GEGLU outputs per chunk
  static constexpr int NCH = J / CH;                       // 40 chunks
  static constexpr int A1_KB_BYTES = 128 * 128;            // one k-block of the row tile
  static constexpr int A1_BYTES = KB1 * A1_KB_BYTES;       // 80 KB
  static constexpr int B1_KB_BYTES = 64 * 128;             // 64 packed rows x one k-block
  static constexpr int B1_BYTES = KB1 * B1_KB_BYTES;       // 40 KB per chunk
  static constexpr int B1_STAGES = 2;                      // ring of whole chunks (one 3-D TMA box each)
  static constexpr int B2_HALF_BYTES = 160 * 64;           // 160 output rows x 32 columns
  static constexpr int B2_BYTES = 2 * B2_HALF_BYTES;       // 20 KB per chunk
  static constexpr int A2_BYTES = 128 * CH * 2;            // 8 KB
  static constexpr int NBAR = 48;
  static constexpr int NB = 3;  // acc1 (TMEM) and A2 (smem) buffers: two chunks of epilogue in flight behind the tensor core
  static constexpr int SMEM_BYTES = A1_BYTES + B1_STAGES * B1_BYTES + 2 * B2_BYTES + NB * A2_BYTES + NBAR * 8 + 16;
  static constexpr int THREADS = 384;
  static constexpr int TMEM_ACC1 = 320;  // first acc1 buffer (64 columns each)
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};

__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <typename T>
__global__ void __launch_bounds__(FfnCfg::THREADS, 1)
ffn_geglu_fused_kernel(const __grid_constant__ FfnMaps maps, const FfnParams p) {
  using Cfg = FfnCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA1 = smem;
  uint8_t* sB1 = sA1 + Cfg::A1_BYTES;                     // [B1_STAGES][B1_BYTES]
  uint8_t* sB2 = sB1 + Cfg::B1_STAGES * Cfg::B1_BYTES;    // [2][B2_BYTES]
  uint8_t* sA2 = sB2 + 2 * Cfg::B2_BYTES;   // [NB][A2_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA2 + Cfg::NB * Cfg::A2_BYTES);
  uint64_t* a1_full = bars;          // [5] k-block of the row tile landed
  uint64_t* a1_empty = bars + 5;     // [5] the last chunk's GEMM1 has read it
  uint64_t* b1_full = bars + 10;     // [B1_STAGES <= 10] a W1f chunk landed
  uint64_t* b1_empty = bars + 20;    // [B1_STAGES <= 10]
  uint64_t* b2_full = bars + 30;     // [2]
  uint64_t* b2_empty = bars + 32;    // [2]
  uint64_t* acc1_full = bars + 34;   // [3] GEMM1 of a chunk complete
  uint64_t* acc1_empty = bars + 37;  // [3] every epilogue warp has pulled its slice into registers
  uint64_t* a2_full = bars + 40;     // [3] every epilogue warp has written its slice of h
  uint64_t* a2_empty = bars + 43;    // [3] GEMM2 of the chunk has read it
  uint64_t* acc2_full = bars + 46;   // all 40 GEMM2 of the row tile complete
  uint64_t* acc2_empty = bars + 47;  // every epilogue warp has drained acc2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.b1);
    tma_prefetch_desc(&maps.b2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 5; ++i) {
        mbar_init(&a1_full[i], 1);
        mbar_init(&a1_empty[i], 1);
      }
      for (int i = 0; i < Cfg::B1_STAGES; ++i) {
        mbar_init(&b1_full[i], 1);
        mbar_init(&b1_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&b2_full[i], 1);
        mbar_init(&b2_empty[i], 1);
      }
      for (int i = 0; i < Cfg::NB; ++i) {
        mbar_init(&acc1_full[i], 1);
        mbar_init(&acc1_empty[i], 8);
        mbar_init(&a2_full[i], 8);
        mbar_init(&a2_empty[i], 1);
      }
      mbar_init(acc2_full, 1);
      mbar_init(acc2_empty, 8);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    // =============================== producer: row tile (A1) and W1f chunks ===============================
    if (elect_one()) {
      uint32_t g = 0;  // chunks issued so far (all row tiles)
      uint32_t st1 = 0, ph1 = 0;  // W1f ring slot / phase
      int t = 0;       // row tiles so far
      for (int mt = blockIdx.x; mt < p.num_m_tiles; mt += gridDim.x, ++t) {
        for (int kb = 0; kb < Cfg::KB1; ++kb) {
          mbar_wait(&a1_empty[kb], (t & 1) ^ 1);
          mbar_expect_tx(&a1_full[kb], Cfg::A1_KB_BYTES);
          tma_load_2d(sA1 + kb * Cfg::A1_KB_BYTES, &maps.a1, &a1_full[kb], kb * 64, mt * 128);
        }
        for (int j = 0; j < Cfg::NCH; ++j, ++g) {
          mbar_wait(&b1_empty[st1], ph1 ^ 1);
          FFN_STAMP(g, 0);
          mbar_expect_tx(&b1_full[st1], Cfg::B1_BYTES);
          tma_load_3d(sB1 + st1 * Cfg::B1_BYTES, &maps.b1, &b1_full[st1], 0, j * 64, 0);
          FFN_STAMP(g, 1);
          if (++st1 == Cfg::B1_STAGES) {
            st1 = 0;
            ph1 ^= 1;
          }
        }
      }
    }
  } else if (warp == 10) {
    // =============================== producer: W2 chunks ===============================
    if (elect_one()) {
      uint32_t g = 0;
      for (int mt = blockIdx.x; mt < p.num_m_tiles; mt += gridDim.x) {
        for (int j = 0; j < Cfg::NCH; ++j, ++g) {
          const int b = g & 1;
          mbar_wait(&b2_empty[b], ((g >> 1) & 1) ^ 1);
          FFN_STAMP(g, 2);
          mbar_expect_tx(&b2_full[b], Cfg::B2_BYTES);
          tma_load_2d(sB2 + b * Cfg::B2_BYTES, &maps.b2, &b2_full[b], j * Cfg::CH, 0);
          tma_load_2d(sB2 + b * Cfg::B2_BYTES + Cfg::B2_HALF_BYTES, &maps.b2, &b2_full[b], j * Cfg::CH, 160);
          FFN_STAMP(g, 3);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer, GEMM1 ===============================
    // GEMM1 and GEMM2 are issued by TWO threads (warps 1 and 11): with one thread, its serial chain per chunk - two barrier
    // waits + 20 MMAs (back-pressured by the tensor core) + two more waits + 4 MMAs - was the chunk period (2 550 cycles
    // against 1 280 of tensor work; timeline trace).  The two streams touch different accumulators and are ordered only
    // through the barriers (tcgen05.commit tracks the issuing thread's own MMAs).
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc_f16(DT<T>::umma_fmt, 128, 64, 0, 0);
      // Descriptors: built once; an operand at byte offset `off` from the base is desc + (off >> 4) (the start-address field
      // is the low 14 bits in units of 16 B and every shared-memory offset here stays below 256 KB: no carry into the
      // next field), so the issue loop costs one 64-bit add per operand.
      const uint64_t a1_desc = umma_smem_desc(smem_u32(sA1), 16, 1024, UMMA_SWIZZLE_128B);
      const uint64_t b1_desc = umma_smem_desc(smem_u32(sB1), 16, 1024, UMMA_SWIZZLE_128B);
      uint32_t g1 = 0, n1 = 0, p1 = 0;  // chunks issued (all row tiles), g1 % NB, (g1 / NB) & 1
      uint32_t st1 = 0, ph1 = 0;        // W1f ring slot / phase
      int t = 0;
      for (int mt = blockIdx.x; mt < p.num_m_tiles; mt += gridDim.x, ++t) {
        for (int j = 0; j < Cfg::NCH; ++j) {
          FFN_STAMP(g1, 4);
          mbar_wait(&acc1_empty[n1], p1 ^ 1);
          FFN_STAMP(g1, 5);
          mbar_wait(&b1_full[st1], ph1);
          FFN_STAMP(g1, 6);
          tc_fence_after();
          const uint32_t d1 = tmem_base + Cfg::TMEM_ACC1 + n1 * 64;
          const uint64_t b1_stage = b1_desc + (uint64_t)((st1 * Cfg::B1_BYTES) >> 4);
#pragma unroll
          for (int kb = 0; kb < Cfg::KB1; ++kb) {
            if (j == 0) {
              mbar_wait(&a1_full[kb], t & 1);
              tc_fence_after();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = a1_desc + (uint64_t)((kb * Cfg::A1_KB_BYTES + k * 32) >> 4);
              const uint64_t bd = b1_stage + (uint64_t)((kb * Cfg::B1_KB_BYTES + k * 32) >> 4);
              umma_f16_ss(d1, ad, bd, idesc1, (kb | k) != 0);
            }
            if (j == Cfg::NCH - 1) umma_commit(&a1_empty[kb]);  // the row tile's k-block may be replaced
          }
          umma_commit(&b1_empty[st1]);
          if (++st1 == Cfg::B1_STAGES) {
            st1 = 0;
            ph1 ^= 1;
          }
          umma_commit(&acc1_full[n1]);
          FFN_STAMP(g1, 7);
          ++g1;
          if (++n1 == Cfg::NB) {
            n1 = 0;
            p1 ^= 1;
          }
        }
      }
    }
  } else if (warp == 11) {
    // =============================== MMA issuer, GEMM2 ===============================
    if (elect_one()) {
      constexpr uint32_t idesc2 = umma_idesc_f16(DT<T>::umma_fmt, 128, 160, 0, 0);
      // A2: no-swizzle K-major [8-col chunk][row][16 B]: chunk pitch 2 KB (LBO), 8-row group pitch 128 B (SBO)
      const uint64_t a2_desc = umma_smem_desc(smem_u32(sA2), 2048, 128, UMMA_SWIZZLE_NONE);
      // W2 chunk half: 160 rows x 64 B, 64B swizzle: 8-row group pitch 512 B, +32 B per 16-element k step
      const uint64_t b2_desc = umma_smem_desc(smem_u32(sB2), 16, 512, UMMA_SWIZZLE_64B);
      uint32_t g2 = 0, n2 = 0, p2 = 0;
      int t = 0;
      for (int mt = blockIdx.x; mt < p.num_m_tiles; mt += gridDim.x, ++t) {
        mbar_wait(acc2_empty, (t & 1) ^ 1);  // the previous row tile's output has been read out of acc2
        tc_fence_after();
        for (int i = 0; i < Cfg::NCH; ++i) {
          const int b = g2 & 1;             // W2 ring stage
          const uint32_t ph = (g2 >> 1) & 1;
          FFN_STAMP(g2, 8);
          mbar_wait(&a2_full[n2], p2);
          FFN_STAMP(g2, 9);
          mbar_wait(&b2_full[b], ph);
          FFN_STAMP(g2, 10);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < Cfg::CH / 16; ++k) {
            const uint64_t ad = a2_desc + (uint64_t)((n2 * Cfg::A2_BYTES + k * 2 * 2048) >> 4);
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
              const uint64_t bd = b2_desc + (uint64_t)((b * Cfg::B2_BYTES + nh * Cfg::B2_HALF_BYTES + k * 32) >> 4);
              umma_f16_ss(tmem_base + nh * 160, ad, bd, idesc2, (i | k) != 0);
            }
          }
          umma_commit(&a2_empty[n2]);
          umma_commit(&b2_empty[b]);
          ++g2;
          if (++n2 == Cfg::NB) {
            n2 = 0;
            p2 ^= 1;
          }
        }
        umma_commit(acc2_full);
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3;            // TMEM lane quarter: rows q*32 .. q*32+31 of the tile
    const int cg = (warp - 2) >> 2;    // column half
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    using T2 = typename DT<T>::T2;
    uint32_t g = 0, nb = 0, pb = 0;  // chunk counter, g % NB, (g / NB) & 1
    int t = 0;
    for (int mt = blockIdx.x; mt < p.num_m_tiles; mt += gridDim.x, ++t) {
      const int row = q * 32 + lane;
      const int m = mt * 128 + row;
      const int mc = min(m, p.M - 1);
      // rstd of this thread's row (the weights are centred, so the mean needs no separate term)
      float sx = 0.f, sxx = 0.f;
      for (int pp = 0; pp < p.stats_parts; ++pp) {
        const float2 v = __ldg(&p.stats_in[(size_t)pp * p.M + mc]);
        sx += v.x;
        sxx += v.y;
      }
      const float inv_k = 1.0f / (float)Cfg::C;
      const float mean = sx * inv_k;
      const float rstd = rsqrtf(fmaxf(sxx * inv_k - mean * mean, 0.f) + p.ln_eps);
      // constant vector of this warp's 16 outputs (h and gate parts) of a chunk: broadcast loads, fetched ONE CHUNK AHEAD
      // (fetched at the top of their own chunk, their L2 latency sat on the epilogue's critical path)
      float4 chn[4], cgn[4];
      auto fetch_c = [&](int j) {
        const float4* ch = reinterpret_cast<const float4*>(p.c1 + (size_t)j * 64 + cg * 16);
        const float4* cgp = reinterpret_cast<const float4*>(p.c1 + (size_t)j * 64 + 32 + cg * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          chn[i] = __ldg(ch + i);
          cgn[i] = __ldg(cgp + i);
        }
      };
      fetch_c(0);
      for (int j = 0; j < Cfg::NCH; ++j, ++g) {
        float chf[16], cgf[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          chf[4 * i] = chn[i].x, chf[4 * i + 1] = chn[i].y, chf[4 * i + 2] = chn[i].z, chf[4 * i + 3] = chn[i].w;
          cgf[4 * i] = cgn[i].x, cgf[4 * i + 1] = cgn[i].y, cgf[4 * i + 2] = cgn[i].z, cgf[4 * i + 3] = cgn[i].w;
        }
        if (j + 1 < Cfg::NCH) fetch_c(j + 1);
        if (warp == 2 && lane == 0) FFN_STAMP(g, 11);
        mbar_wait(&acc1_full[nb], pb);
        if (warp == 2 && lane == 0) FFN_STAMP(g, 12);
        tc_fence_after();
        uint32_t rh[16], rg[16];
        const uint32_t ta = tmem_base + lane_sel + Cfg::TMEM_ACC1 + nb * 64;
        tmem_ld16(ta + cg * 16, rh);
        tmem_ld16(ta + 32 + cg * 16, rg);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[nb]);  // the accumulator lives in registers now
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          v[i] = fmaf(__uint_as_float(rh[i]), rstd, chf[i]) * gelu_erf_f(fmaf(__uint_as_float(rg[i]), rstd, cgf[i]));
        const uint4 p0 = pack8<T>(v), p1 = pack8<T>(v + 8);
        if (warp == 2 && lane == 0) FFN_STAMP(g, 13);
        mbar_wait(&a2_empty[nb], pb ^ 1);  // GEMM2 of the chunk that used this buffer before has read it
        if (warp == 2 && lane == 0) FFN_STAMP(g, 14);
        uint8_t* a2 = sA2 + nb * Cfg::A2_BYTES;
        *reinterpret_cast<uint4*>(a2 + (2 * cg) * 2048 + row * 16) = p0;      // columns 16 cg .. +7
        *reinterpret_cast<uint4*>(a2 + (2 * cg + 1) * 2048 + row * 16) = p1;  // columns 16 cg + 8 .. +15
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[nb]);
        if (warp == 2 && lane == 0) FFN_STAMP(g, 15);
        if (++nb == Cfg::NB) {
          nb = 0;
          pb ^= 1;
        }
      }
      // ---- output of the row tile: y = round(acc2 + b2) + y
      mbar_wait(acc2_full, t & 1);
      tc_fence_after();
      const T* res = reinterpret_cast<const T*>(p.res) + (size_t)mc * Cfg::C + cg * 160;
      T* out = reinterpret_cast<T*>(p.out) + (size_t)mc * Cfg::C + cg * 160;
#pragma unroll 1
      for (int c = 0; c < 160; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + lane_sel + cg * 160 + c, r);
        const uint4 rv0 = *reinterpret_cast<const uint4*>(res + c);      // (plain loads: res may alias out)
        const uint4 rv1 = *reinterpret_cast<const uint4*>(res + c + 8);
        const float4* bp = reinterpret_cast<const float4*>(p.bias2 + cg * 160 + c);
        float bv[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x = __ldg(bp + i);
          bv[4 * i] = x.x, bv[4 * i + 1] = x.y, bv[4 * i + 2] = x.z, bv[4 * i + 3] = x.w;
        }
        tmem_wait_ld();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + bv[i];
        uint4 o0 = pack8<T>(v), o1 = pack8<T>(v + 8);
        T2* a0 = reinterpret_cast<T2*>(&o0);
        T2* a1 = reinterpret_cast<T2*>(&o1);
        const T2* b0 = reinterpret_cast<const T2*>(&rv0);
        const T2* b1 = reinterpret_cast<const T2*>(&rv1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a0[i] = __hadd2(a0[i], b0[i]);
          a1[i] = __hadd2(a1[i], b1[i]);
        }
        if (m < p.M) {
          *reinterpret_cast<uint4*>(out + c) = o0;
          *reinterpret_cast<uint4*>(out + c + 8) = o1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------------------------------
// CTA-pair version (cluster of 2, tcgen05 cta_group::2, M = 256): the pair owns 256 rows; CTA r keeps rows [128 r, +128)
// resident and holds HALF of every weight chunk (W1f: 32 of the chunk's 64 packed rows; W2: 80 of each 160-row half), so the
// same shared memory gives the W1f ring four stages instead of two (the refill latency of ~2 100 cycles was what bound the
// single-CTA kernel) and every SM pulls half the weight bytes.  MMAs are issued by the leader CTA's two issuer threads;
// loads of both CTAs complete on the leader's barriers, commits are multicast to both CTAs, and the epilogue warps of both
// CTAs arrive on the leader's barriers (same scheme as the CTA-pair GEMM in gemm_tcgen05.cuh).
// ------------------------------------------------------------------------------------------------------------------------
struct FfnPairCfg {
  static constexpr int C = 320, J = 4 * C, KB1 = C / 64, CH = 32, NCH = J / CH;
  static constexpr int A1_KB_BYTES = 128 * 128;
  static constexpr int A1_BYTES = KB1 * A1_KB_BYTES;        // 80 KB
  static constexpr int B1_KB_BYTES = 32 * 128;              // this CTA's 32 packed rows x one k-block
  static constexpr int B1_BYTES = KB1 * B1_KB_BYTES;        // 20 KB per chunk
  static constexpr int B1_STAGES = 4;
  static constexpr int B2_HALF_BYTES = 80 * 64;             // this CTA's 80 rows of one 160-row half x 32 columns
  static constexpr int B2_BYTES = 2 * B2_HALF_BYTES;        // 10 KB per chunk
  static constexpr int B2_STAGES = 3;
  static constexpr int A2_BYTES = 128 * CH * 2;             // 8 KB
  static constexpr int NB = 3;
  static constexpr int NBAR = 48;
  static constexpr int SMEM_BYTES = A1_BYTES + B1_STAGES * B1_BYTES + B2_STAGES * B2_BYTES + NB * A2_BYTES + NBAR * 8 + 16;
  static constexpr int THREADS = 384;
  static constexpr int TMEM_ACC1 = 320;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};

__device__ __forceinline__ void tma_load_3d_2sm(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// maps: a1 as in the single-CTA kernel; b1 box (64, 32, C/64); b2 box (32, 80)
template <typename T>
__global__ void __launch_bounds__(FfnPairCfg::THREADS, 1)
ffn_geglu_fused_pair_kernel(const __grid_constant__ FfnMaps maps, const FfnParams p) {
  using Cfg = FfnPairCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA1 = smem;
  uint8_t* sB1 = sA1 + Cfg::A1_BYTES;
  uint8_t* sB2 = sB1 + Cfg::B1_STAGES * Cfg::B1_BYTES;
  uint8_t* sA2 = sB2 + Cfg::B2_STAGES * Cfg::B2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA2 + Cfg::NB * Cfg::A2_BYTES);
  uint64_t* a1_full = bars;          // [5]  (leader's is used)
  uint64_t* a1_empty = bars + 5;     // [5]  multicast commit
  uint64_t* b1_full = bars + 10;     // [4]  leader
  uint64_t* b1_empty = bars + 14;    // [4]  multicast commit
  uint64_t* b2_full = bars + 18;     // [3]  leader
  uint64_t* b2_empty = bars + 21;    // [3]  multicast commit
  uint64_t* acc1_full = bars + 24;   // [3]  multicast commit
  uint64_t* acc1_empty = bars + 27;  // [3]  leader, 16 arrivals
  uint64_t* a2_full = bars + 30;     // [3]  leader, 16 arrivals
  uint64_t* a2_empty = bars + 33;    // [3]  multicast commit
  uint64_t* acc2_full = bars + 36;   //      multicast commit
  uint64_t* acc2_empty = bars + 37;  //      leader, 16 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int wid = (int)(blockIdx.x >> 1), nworkers = (int)(gridDim.x >> 1);
  const int num_pairs = (p.num_m_tiles + 1) / 2;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.b1);
    tma_prefetch_desc(&maps.b2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 5; ++i) {
        mbar_init(&a1_full[i], 1);
        mbar_init(&a1_empty[i], 1);
      }
      for (int i = 0; i < Cfg::B1_STAGES; ++i) {
        mbar_init(&b1_full[i], 1);
        mbar_init(&b1_empty[i], 1);
      }
      for (int i = 0; i < Cfg::B2_STAGES; ++i) {
        mbar_init(&b2_full[i], 1);
        mbar_init(&b2_empty[i], 1);
      }
      for (int i = 0; i < Cfg::NB; ++i) {
        mbar_init(&acc1_full[i], 1);
        mbar_init(&acc1_empty[i], 16);
        mbar_init(&a2_full[i], 16);
        mbar_init(&a2_empty[i], 1);
      }
      mbar_init(acc2_full, 1);
      mbar_init(acc2_empty, 16);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta<512>(tmem_slot);
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();
  // arrival on the LEADER's copy of a barrier (epilogue warps of both CTAs)
  auto arrive_leader = [&](uint64_t* bar) { mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0)); };

  if (warp == 0) {
    // =============================== producer: row tile (A1) and this CTA's half of the W1f chunks ===============================
    if (elect_one()) {
      uint32_t st1 = 0, ph1 = 0;
      int t = 0;
      for (int pt = wid; pt < num_pairs; pt += nworkers, ++t) {
        const int mt = 2 * pt + (int)rank;
        for (int kb = 0; kb < Cfg::KB1; ++kb) {
          mbar_wait(&a1_empty[kb], (t & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&a1_full[kb], 2 * Cfg::A1_KB_BYTES);
          tma_load_2d_2sm(sA1 + kb * Cfg::A1_KB_BYTES, &maps.a1, &a1_full[kb], kb * 64, mt * 128);
        }
        for (int j = 0; j < Cfg::NCH; ++j) {
          mbar_wait(&b1_empty[st1], ph1 ^ 1);
          if (rank == 0) mbar_expect_tx(&b1_full[st1], 2 * Cfg::B1_BYTES);
          tma_load_3d_2sm(sB1 + st1 * Cfg::B1_BYTES, &maps.b1, &b1_full[st1], 0, j * 64 + (int)rank * 32, 0);
          if (++st1 == Cfg::B1_STAGES) {
            st1 = 0;
            ph1 ^= 1;
          }
        }
      }
    }
  } else if (warp == 10) {
    // =============================== producer: this CTA's half of the W2 chunks ===============================
    if (elect_one()) {
      uint32_t st2 = 0, ph2 = 0;
      for (int pt = wid; pt < num_pairs; pt += nworkers) {
        for (int j = 0; j < Cfg::NCH; ++j) {
          mbar_wait(&b2_empty[st2], ph2 ^ 1);
          if (rank == 0) mbar_expect_tx(&b2_full[st2], 2 * Cfg::B2_BYTES);
          uint8_t* dst = sB2 + st2 * Cfg::B2_BYTES;
          tma_load_2d_2sm(dst, &maps.b2, &b2_full[st2], j * Cfg::CH, (int)rank * 80);
          tma_load_2d_2sm(dst + Cfg::B2_HALF_BYTES, &maps.b2, &b2_full[st2], j * Cfg::CH, 160 + (int)rank * 80);
          if (++st2 == Cfg::B2_STAGES) {
            st2 = 0;
            ph2 ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer, GEMM1 (leader CTA only) ===============================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc_f16(DT<T>::umma_fmt, 256, 64, 0, 0);
      const uint64_t a1_desc = umma_smem_desc(smem_u32(sA1), 16, 1024, UMMA_SWIZZLE_128B);
      const uint64_t b1_desc = umma_smem_desc(smem_u32(sB1), 16, 1024, UMMA_SWIZZLE_128B);
      uint32_t n1 = 0, p1 = 0, st1 = 0, ph1 = 0;
      int t = 0;
      for (int pt = wid; pt < num_pairs; pt += nworkers, ++t) {
        for (int j = 0; j < Cfg::NCH; ++j) {
          mbar_wait(&acc1_empty[n1], p1 ^ 1);
          mbar_wait(&b1_full[st1], ph1);
          tc_fence_after();
          const uint32_t d1 = tmem_base + Cfg::TMEM_ACC1 + n1 * 64;
          const uint64_t b1_stage = b1_desc + (uint64_t)((st1 * Cfg::B1_BYTES) >> 4);
#pragma unroll
          for (int kb = 0; kb < Cfg::KB1; ++kb) {
            if (j == 0) {
              mbar_wait(&a1_full[kb], t & 1);
              tc_fence_after();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = a1_desc + (uint64_t)((kb * Cfg::A1_KB_BYTES + k * 32) >> 4);
              const uint64_t bd = b1_stage + (uint64_t)((kb * Cfg::B1_KB_BYTES + k * 32) >> 4);
              umma_f16_ss_2cta(d1, ad, bd, idesc1, (kb | k) != 0);
            }
            if (j == Cfg::NCH - 1) umma_commit_2cta(&a1_empty[kb]);
          }
          umma_commit_2cta(&b1_empty[st1]);
          umma_commit_2cta(&acc1_full[n1]);
          if (++st1 == Cfg::B1_STAGES) {
            st1 = 0;
            ph1 ^= 1;
          }
          if (++n1 == Cfg::NB) {
            n1 = 0;
            p1 ^= 1;
          }
        }
      }
    }
  } else if (warp == 11) {
    // =============================== MMA issuer, GEMM2 (leader CTA only) ===============================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc2 = umma_idesc_f16(DT<T>::umma_fmt, 256, 160, 0, 0);
      const uint64_t a2_desc = umma_smem_desc(smem_u32(sA2), 2048, 128, UMMA_SWIZZLE_NONE);
      const uint64_t b2_desc = umma_smem_desc(smem_u32(sB2), 16, 512, UMMA_SWIZZLE_64B);
      uint32_t n2 = 0, p2 = 0, st2 = 0, ph2 = 0;
      int t = 0;
      for (int pt = wid; pt < num_pairs; pt += nworkers, ++t) {
        mbar_wait(acc2_empty, (t & 1) ^ 1);
        tc_fence_after();
        for (int i = 0; i < Cfg::NCH; ++i) {
          mbar_wait(&a2_full[n2], p2);
          mbar_wait(&b2_full[st2], ph2);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < Cfg::CH / 16; ++k) {
            const uint64_t ad = a2_desc + (uint64_t)((n2 * Cfg::A2_BYTES + k * 2 * 2048) >> 4);
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
              const uint64_t bd = b2_desc + (uint64_t)((st2 * Cfg::B2_BYTES + nh * Cfg::B2_HALF_BYTES + k * 32) >> 4);
              umma_f16_ss_2cta(tmem_base + nh * 160, ad, bd, idesc2, (i | k) != 0);
            }
          }
          umma_commit_2cta(&a2_empty[n2]);
          umma_commit_2cta(&b2_empty[st2]);
          if (++st2 == Cfg::B2_STAGES) {
            st2 = 0;
            ph2 ^= 1;
          }
          if (++n2 == Cfg::NB) {
            n2 = 0;
            p2 ^= 1;
          }
        }
        umma_commit_2cta(acc2_full);
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // =============================== epilogue warps (both CTAs) ===============================
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    using T2 = typename DT<T>::T2;
    uint32_t nb = 0, pb = 0;
    int t = 0;
    for (int pt = wid; pt < num_pairs; pt += nworkers, ++t) {
      const int mt = 2 * pt + (int)rank;
      const int row = q * 32 + lane;
      const int m = mt * 128 + row;
      const int mc = min(m, p.M - 1);
      float sx = 0.f, sxx = 0.f;
      for (int pp = 0; pp < p.stats_parts; ++pp) {
        const float2 v = __ldg(&p.stats_in[(size_t)pp * p.M + mc]);
        sx += v.x;
        sxx += v.y;
      }
      const float inv_k = 1.0f / (float)Cfg::C;
      const float mean = sx * inv_k;
      const float rstd = rsqrtf(fmaxf(sxx * inv_k - mean * mean, 0.f) + p.ln_eps);
      float4 chn[4], cgn[4];
      auto fetch_c = [&](int j) {
        const float4* ch = reinterpret_cast<const float4*>(p.c1 + (size_t)j * 64 + cg * 16);
        const float4* cgp = reinterpret_cast<const float4*>(p.c1 + (size_t)j * 64 + 32 + cg * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          chn[i] = __ldg(ch + i);
          cgn[i] = __ldg(cgp + i);
        }
      };
      fetch_c(0);
      for (int j = 0; j < Cfg::NCH; ++j) {
        float chf[16], cgf[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          chf[4 * i] = chn[i].x, chf[4 * i + 1] = chn[i].y, chf[4 * i + 2] = chn[i].z, chf[4 * i + 3] = chn[i].w;
          cgf[4 * i] = cgn[i].x, cgf[4 * i + 1] = cgn[i].y, cgf[4 * i + 2] = cgn[i].z, cgf[4 * i + 3] = cgn[i].w;
        }
        if (j + 1 < Cfg::NCH) fetch_c(j + 1);
        mbar_wait(&acc1_full[nb], pb);
        tc_fence_after();
        uint32_t rh[16], rg[16];
        const uint32_t ta = tmem_base + lane_sel + Cfg::TMEM_ACC1 + nb * 64;
        tmem_ld16(ta + cg * 16, rh);
        tmem_ld16(ta + 32 + cg * 16, rg);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&acc1_empty[nb]);
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          v[i] = fmaf(__uint_as_float(rh[i]), rstd, chf[i]) * gelu_erf_f(fmaf(__uint_as_float(rg[i]), rstd, cgf[i]));
        const uint4 p0 = pack8<T>(v), p1 = pack8<T>(v + 8);
        mbar_wait(&a2_empty[nb], pb ^ 1);
        uint8_t* a2 = sA2 + nb * Cfg::A2_BYTES;
        *reinterpret_cast<uint4*>(a2 + (2 * cg) * 2048 + row * 16) = p0;
        *reinterpret_cast<uint4*>(a2 + (2 * cg + 1) * 2048 + row * 16) = p1;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive_leader(&a2_full[nb]);
        if (++nb == Cfg::NB) {
          nb = 0;
          pb ^= 1;
        }
      }
      mbar_wait(acc2_full, t & 1);
      tc_fence_after();
      const T* res = reinterpret_cast<const T*>(p.res) + (size_t)mc * Cfg::C + cg * 160;
      T* out = reinterpret_cast<T*>(p.out) + (size_t)mc * Cfg::C + cg * 160;
#pragma unroll 1
      for (int c = 0; c < 160; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + lane_sel + cg * 160 + c, r);
        const uint4 rv0 = *reinterpret_cast<const uint4*>(res + c);
        const uint4 rv1 = *reinterpret_cast<const uint4*>(res + c + 8);
        const float4* bp = reinterpret_cast<const float4*>(p.bias2 + cg * 160 + c);
        float bv[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x = __ldg(bp + i);
          bv[4 * i] = x.x, bv[4 * i + 1] = x.y, bv[4 * i + 2] = x.z, bv[4 * i + 3] = x.w;
        }
        tmem_wait_ld();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + bv[i];
        uint4 o0 = pack8<T>(v), o1 = pack8<T>(v + 8);
        T2* a0 = reinterpret_cast<T2*>(&o0);
        T2* a1 = reinterpret_cast<T2*>(&o1);
        const T2* b0 = reinterpret_cast<const T2*>(&rv0);
        const T2* b1 = reinterpret_cast<const T2*>(&rv1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a0[i] = __hadd2(a0[i], b0[i]);
          a1[i] = __hadd2(a1[i], b1[i]);
        }
        if (m < p.M) {
          *reinterpret_cast<uint4*>(out + c) = o0;
          *reinterpret_cast<uint4*>(out + c + 8) = o1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(acc2_empty);
    }
  }
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();  // neither CTA may exit while its peer can still touch its shared memory / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

}  // namespace rcdm
