// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M, N] = sum over K-segments of A_seg[M, K_seg] * W[N, K_total]^T   (+ epilogue)
//
// Persistent kernel, one CTA per SM, looping over 128 x BN output tiles.  Warp roles (192 threads):
//   warp 0     TMA producer  (one elected lane)  : HBM/L2 -> 128B-swizzled smem ring (5-6 stages), runs ahead
//                                                  across tile boundaries
//   warp 1     MMA issuer    (one elected lane)  : tcgen05.mma into one of TWO TMEM accumulators; owns TMEM alloc
//   warps 2-9  epilogue (8 warps)                : residual prefetch (coalesced, before the accumulator is ready)
//                                                  -> tcgen05.ld (+bias / GEGLU) -> 16-bit smem staging slab
//                                                  -> 16-byte lane-contiguous global stores (+residual)
// so the main loop of tile i+1 overlaps the epilogue of tile i.
//
// The K loop runs over up to three "segments", each reading A through its own TMA tensor map(s):
//   SEG_PLAIN    A is a row-major [M, K] matrix             (Linear, conv1x1, im2col'ed conv_in)
//   SEG_CONV3    A is an NHWC activation; 3x3, stride 1, pad 1: for each tap the SAME 4-D tensor map is
//                read at spatial offset (kx-1, ky-1); TMA out-of-bounds zero fill implements the padding,
//                so no im2col buffer ever exists ("im2col-free" implicit GEMM)
//   SEG_CONV3S2  3x3, stride 2, pad 1: four parity-subsampled 4-D maps (y%2, x%2), tap -> (map, offset)
// Segments accumulate into the same TMEM tile, which is how conv2 + the 1x1 shortcut of a resblock (whose input
// is the un-materialised concat [h | skip]) become a single launch.
#pragma once
#include "common.cuh"

namespace rcdm {

enum : int { SEG_PLAIN = 0, SEG_CONV3 = 1, SEG_CONV3S2 = 2 };

struct GemmSeg {
  int mode;     // SEG_*
  int tmap;     // index of the first A tensor map of this segment
  int cblocks;  // 64-wide K blocks (per tap for conv segments)
};

struct GemmParams {
  int M, N;       // rows, accumulator columns (for GEGLU: 2x the output width)
  int num_kb;     // total number of 64-wide K blocks over all segments
  int nseg;
  GemmSeg seg[3];
  // conv tile geometry: a 128-row M tile is a (tn images) x (th rows) x (tw cols) box of the OUTPUT grid
  int tw, th, tn, tiles_x, tiles_y;
  int num_m_tiles, num_n_tiles;  // persistent tile loop: tile = m_tile * num_n_tiles + n_tile
  // epilogue
  void* out;
  int ldo;
  const float* bias;  // [N] (GEGLU: packed like the weight rows) or nullptr
  const void* res;    // residual [M, ldr] or nullptr (may alias out)
  int ldr;
  int geglu;          // 1: out[m, j] = (acc[j] + b[j]) * gelu(acc[BN/2 + j] + b[BN/2 + j]) per tile
};

struct GemmMaps {
  CUtensorMap a[4];
  CUtensorMap b;
};

template <int BN> struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int STAGES = (BN <= 64) ? 6 : 5;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // 8 epilogue warps, each with a private slab: 32 rows x (BN/2 cols * 2 B + 16 B pad)
  static constexpr int SLAB_BYTES = 32 * (BN + 16);
  static constexpr int STAGING_BYTES = 8 * SLAB_BYTES;
  static constexpr int ACC_STRIDE = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // TMEM columns between the 2 accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int BIAS_BYTES = 2 * BN * 4;  // two buffers (tile parity) of BN fp32 bias values
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + BIAS_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};

// Persistent, warp-specialised: grid = min(#tiles, #SMs), one CTA per SM.  Tiles are visited round-robin
// (n fastest, so CTAs running concurrently share A rows in L2).  The TMA producer runs ahead across tile
// boundaries; two TMEM accumulators let the MMA warp start tile i+1 while the epilogue warps drain tile i.
template <typename T, int BN>
__global__ void __launch_bounds__(320, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* bias_sm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.a[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full_bar[i], 1);
        mbar_init(&tmem_empty_bar[i], 256);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.num_n_tiles, m_tile = tile / p.num_n_tiles;
        int t = m_tile;  // conv tile origin (only used by conv segments)
        const int tx = t % p.tiles_x;
        t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int tb = t / p.tiles_y;
        const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tb * p.tn;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg sg = p.seg[s];
          const int ntap = (sg.mode == SEG_PLAIN) ? 1 : 9;
          for (int tap = 0; tap < ntap; ++tap) {
            int mi = sg.tmap, dx = 0, dy = 0;
            if (sg.mode == SEG_CONV3) {
              dy = tap / 3 - 1;
              dx = tap % 3 - 1;
            } else if (sg.mode == SEG_CONV3S2) {
              const int ky = tap / 3, kx = tap % 3;
              const int py = (ky + 1) & 1, px = (kx + 1) & 1;  // parity of (2y + ky - 1)
              dy = (ky == 0) ? -1 : 0;
              dx = (kx == 0) ? -1 : 0;
              mi = sg.tmap + py * 2 + px;
            }
            for (int c = 0; c < sg.cblocks; ++c, ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              void* sa = smem_a + stage * Cfg::A_BYTES;
              void* sb = smem_b + stage * Cfg::B_BYTES;
              if (sg.mode == SEG_PLAIN)
                tma_load_2d(sa, &maps.a[mi], &full_bar[stage], c * 64, m_tile * 128);
              else
                tma_load_4d(sa, &maps.a[mi], &full_bar[stage], c * 64, x0 + dx, y0 + dy, n0);
              tma_load_2d(sb, &maps.b, &full_bar[stage], kb * 64, n_tile * BN);
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =================================== MMA issuer ===================================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(DT<T>::umma_fmt, 128, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // K-major, 128B swizzle: rows at 128 B, 8-row groups at 1024 B; +32 B per 16-element K step
            const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
            const uint64_t bd = umma_smem_desc(b_addr + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
            umma_f16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // =================================== epilogue (8 warps) ===================================
    // warp -> (q, hs): q = TMEM lane quarter (rows q*32..+31), hs = which half of the tile's output columns.
    // phase 1 (thread <-> row):  TMEM -> registers, + bias (smem broadcast), GEGLU -> 16-bit private slab
    // phase 2 (lane <-> fixed 16-byte column chunk, RPI rows per pass): slab (+ prefetched residual, packed
    //          half2 add == fp32 add + one rounding) -> coalesced global stores
    const int q = warp & 3;
    const int hs = (warp - 2) >> 2;
    T* out = reinterpret_cast<T*>(p.out);
    const T* res = reinterpret_cast<const T*>(p.res);
    const bool vec_ok = (p.N % 8 == 0) && (p.ldo % 8 == 0) && (!p.res || p.ldr % 8 == 0);
    constexpr int OUT_W = BN;            // accumulator columns per tile
    constexpr int HALF = BN / 2;         // accumulator columns per warp (plain) ...
    constexpr int W_COLS = HALF;         // max output columns per warp (GEGLU uses HALF / 2)
    constexpr int PITCH = W_COLS * 2 + 16;
    uint8_t* slab = staging + (size_t)(warp - 2) * Cfg::SLAB_BYTES;
    const int n_total = p.geglu ? p.N / 2 : p.N;
    const int wcols = p.geglu ? HALF / 2 : HALF;   // output columns this warp produces per tile
    const int CH = wcols / 8;                      // 16-byte chunks per slab row
    const int RPI = 32 / CH;                       // rows per phase-2 pass
    const int l_row = lane / CH, l_chunk = lane - l_row * CH;
    const bool l_active = l_row < RPI;
    constexpr int MAX_PASS = (BN == 160) ? 11 : 8;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % p.num_n_tiles, m_tile = tile / p.num_n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + acc * Cfg::ACC_STRIDE + (uint32_t(q * 32) << 16);
      const int m_warp = m_tile * 128 + q * 32;
      if (vec_ok) {
        const int n_warp = n_tile * (p.geglu ? HALF : OUT_W) + hs * wcols;  // first output column of this warp
        const bool chunk_ok = l_active && (n_warp + l_chunk * 8 < n_total);
        // ---- residual prefetch (coalesced; issued before the accumulator is ready => hidden by the main loop)
        uint4 rv[MAX_PASS];
        if (res) {
#pragma unroll
          for (int u = 0; u < MAX_PASS; ++u) {
            const int rr = u * RPI + l_row;
            if (chunk_ok && rr < 32 && m_warp + rr < p.M)
              rv[u] = *reinterpret_cast<const uint4*>(res + (size_t)(m_warp + rr) * p.ldr + n_warp + l_chunk * 8);
          }
        }
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        // ---- bias slice of this warp's column half -> smem[tile parity].  The four warps sharing `hs` write
        // identical values (benign); filling AFTER the wait makes the parity double-buffer race-free (tile i+2
        // cannot become ready before every thread finished phase 1 of tile i).
        float* bsm = bias_sm + (it & 1) * BN + hs * HALF;
        if (p.bias) {
          if (!p.geglu) {
            for (int i = lane; i < HALF; i += 32) bsm[i] = (n_warp + i < n_total) ? __ldg(p.bias + n_warp + i) : 0.f;
          } else {  // packed GEGLU bias: tile-local [h (HALF) | gate (HALF)]; h/gate columns hs*wcols..+wcols
            for (int i = lane; i < wcols; i += 32) {
              bsm[i] = __ldg(p.bias + n_tile * BN + hs * wcols + i);
              bsm[wcols + i] = __ldg(p.bias + n_tile * BN + HALF + hs * wcols + i);
            }
          }
        } else {
          for (int i = lane; i < HALF; i += 32) bsm[i] = 0.f;
        }
        __syncwarp();
        // ---- phase 1
        if (!p.geglu) {
#pragma unroll 1
          for (int c = 0; c < HALF; c += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + hs * HALF + c, r);
            tmem_wait_ld();
            float v[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 b4 = *reinterpret_cast<const float4*>(bsm + c + g * 4);  // smem broadcast
              v[g * 4] = __uint_as_float(r[g * 4]) + b4.x;
              v[g * 4 + 1] = __uint_as_float(r[g * 4 + 1]) + b4.y;
              v[g * 4 + 2] = __uint_as_float(r[g * 4 + 2]) + b4.z;
              v[g * 4 + 3] = __uint_as_float(r[g * 4 + 3]) + b4.w;
            }
            *reinterpret_cast<uint4*>(slab + lane * PITCH + c * 2) = pack8<T>(v);
            *reinterpret_cast<uint4*>(slab + lane * PITCH + c * 2 + 16) = pack8<T>(v + 8);
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < HALF / 2; c += 16) {
            uint32_t rh[16], rg[16];
            tmem_ld16(taddr + hs * (HALF / 2) + c, rh);
            tmem_ld16(taddr + HALF + hs * (HALF / 2) + c, rg);
            tmem_wait_ld();
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              v[i] = (__uint_as_float(rh[i]) + bsm[c + i]) * gelu_erf_f(__uint_as_float(rg[i]) + bsm[wcols + c + i]);
            *reinterpret_cast<uint4*>(slab + lane * PITCH + c * 2) = pack8<T>(v);
            *reinterpret_cast<uint4*>(slab + lane * PITCH + c * 2 + 16) = pack8<T>(v + 8);
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);  // accumulator drained: the MMA warp may reuse it
        __syncwarp();
        // ---- phase 2
        if (chunk_ok) {
#pragma unroll
          for (int u = 0; u < MAX_PASS; ++u) {
            const int rr = u * RPI + l_row;
            if (rr < 32 && m_warp + rr < p.M) {
              uint4 sv = *reinterpret_cast<const uint4*>(slab + rr * PITCH + l_chunk * 16);
              if (res) {
                using T2 = typename DT<T>::T2;
                T2* a2 = reinterpret_cast<T2*>(&sv);
                const T2* b2 = reinterpret_cast<const T2*>(&rv[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) a2[i] = __hadd2(a2[i], b2[i]);
              }
              *reinterpret_cast<uint4*>(out + (size_t)(m_warp + rr) * p.ldo + n_warp + l_chunk * 8) = sv;
            }
          }
        }
        __syncwarp();  // the slab is rewritten for the next tile
      } else {
        // ---- scalar fallback (conv_out: N = 4): thread <-> row, direct stores; only the hs == 0 warps work
        const int m = m_warp + lane;
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        if (hs == 0) {
#pragma unroll 1
          for (int c = 0; c < BN; c += 16) {
            const int n = n_tile * BN + c;
            if (n >= p.N) break;  // warp-uniform
            uint32_t r[16];
            tmem_ld16(taddr + c, r);
            tmem_wait_ld();
            if (m < p.M) {
              for (int i = 0; i < 16; ++i) {
                if (n + i < p.N) {
                  float x = __uint_as_float(r[i]);
                  if (p.bias) x += __ldg(p.bias + n + i);
                  if (res) x += DT<T>::to_f(res[(size_t)m * p.ldr + n + i]);
                  out[(size_t)m * p.ldo + n + i] = DT<T>::from_f(x);
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace rcdm
