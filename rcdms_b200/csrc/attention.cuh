// Attention kernels for sm_100a.
//
// flash_attn4_kernel: softmax(scale * Q K^T) V without materialising the scores (the reference writes an
// (80, 4096, 4096) score tensor per 64x64 layer: attention.py:170-199).  One CTA = 128 queries of one
// (image, head).  Both GEMMs run on tcgen05 with accumulators in TMEM:
//     S_j = Q K_j^T (128 x 64 fp32, one TMEM buffer)         O += P_j V_j   (128 x DPAD fp32, TMEM cols [64, ...))
// Q/K/V head slices are fetched straight out of the fused projection output [rows, ld] by 5-D TMA maps
// (8 elems, row, 16-byte chunk, head, image) into the no-swizzle "interleaved" UMMA layout
// [chunk][row][8 elems]; out-of-range chunks / rows are zero-filled by TMA, which pads head_dim 40 -> 48 and
// ragged KV lengths (context L = 85 / 91) for free.  V is consumed as an MN-major B operand, so no transpose.
// Warps 0-3: online softmax (thread <-> query row), P written to smem as the A operand of the second GEMM.
// Warp 4 (one lane): MMA issuer.  Warp 5 (one lane): TMA loads.
//
// temporal_attn_kernel: the motion modules' attention over the f = 5 frames at each spatial location
// (motion_module.py:294-354): a 5x5 problem per (location, head) -> CUDA cores, one thread per (location, head),
// q/k/v read once directly from the (b f hw)-ordered token matrix (no "(b f) d c -> (b d) f c" copies).
#pragma once
#include <type_traits>

// Compile-time experiment switch for bottleneck analysis (never set in a product build):
//   1 = no exp2 (FFMA result used), 2 = no TMEM loads of S, 3 = no P stores to smem, 4 = no P V MMAs
#ifndef RCDM_ATTN_EXPERIMENT
#define RCDM_ATTN_EXPERIMENT 0
#endif

#include "common.cuh"

namespace rcdm {

struct AttnParams {
  int S_q, S_kv;  // rows per image for queries / keys
  int heads, d;   // head dim (multiple of 8)
  int batch;      // images
  void* out;      // [batch * S_q, ldo]; head h -> columns [h*d, h*d + d)
  int ldo;
  float scale_log2;  // d^-0.5 * log2(e)
};

struct AttnMaps {
  CUtensorMap q, k, v;
};

// ------------------------------------------------------------------------------------------
// flash attention (5-D TMA head slices, tcgen05 S = Q K^T and O += P V with TMEM accumulators, lazy-rescale online
// softmax, PV-MMA row sums), balanced for the unit that bounds head_dim 40: the MUFU exp2 pipe (ncu of the first
// version, S double-buffered in TMEM and 2 CTAs per SM: XU 58 % busy, tensor 21 %, 8 softmax warps per SM).
//   * each softmax thread pulls its whole 64-column S row into registers and releases the TMEM S buffer at once
//     (s_free), so ONE S buffer suffices: TMEM = 64 (S) + DPAD (O) columns -> 128 columns for DPAD <= 64
//   * with 128 TMEM columns, 2 K/V stages and ~70 KB of shared memory, THREE CTAs share an SM (12 softmax warps
//     instead of 8) and their exp / TMEM-load / barrier phases interleave on the MUFU pipe
//   * the rare re-reference pass works from the registers (no second TMEM read of S)
//   * K and V travel through SEPARATE two-stage rings: the K stage of tile j is free as soon as Q K_j^T has completed
//     (start of the softmax of tile j), so K_{j+2} is requested a whole tile earlier than a joint K/V stage (free only
//     after P_j V_j) allows, and V_{j+1} two tiles before P_{j+1} V_{j+1} needs it.  (ncu with joint stages: the MMA warp
//     spent 40 % of its time waiting for the 16-byte-granular head-slice gathers; a third joint stage cost the third
//     resident CTA its shared memory and was 18 % slower.)
//   * head_dim 40: the V box covers only the 5 real 16-byte chunks; the padded 6th chunk of every V stage is written once
//     at kernel start (a one in column 40, zeros behind it), so the P V MMA accumulates the row sums in O's column 40
//     without any per-tile write
// ------------------------------------------------------------------------------------------
// 2^x on the FMA / ALU pipes (no MUFU): Cody-Waite split x = floor(x) + f by a round-down add of 1.5 * 2^23 (floor(x)
// lands in the low mantissa bits), degree-3 minimax polynomial of 2^f on [0, 1) (max relative error 7.5e-5, a third of
// fp16's half-ulp), exponent patched in with one integer add.  7 FMA/ALU instructions against one quarter-rate MUFU.EX2:
// the d = 40 softmax is bound by the 16 exp/clk/SM of the MUFU, so a fraction of every row's exponentials
// (ATTN_POLY_PER8 of each 8) goes through here and the two pipes run side by side (the FlashAttention-4 trick).
#ifndef RCDM_ATTN_POLY_PER8
#define RCDM_ATTN_POLY_PER8 3
#endif
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = __fadd_rd(x, 12582912.0f);
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.07802393287420273f, 0.22606699168682098f);
  p = fmaf(p, f, 0.6958341598510742f);
  p = fmaf(p, f, 0.9999250769615173f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// Timeline instrumentation (variant builds only, -DRCDM_ATTN_TRACE=1; scripts/attn_trace.py): clock64 stamps of the barrier
// hand-offs of every (CTA, K/V tile) into a device array, read back through rcdm_debug_attn_trace_read.
#ifndef RCDM_ATTN_TRACE
#define RCDM_ATTN_TRACE 0
#endif
#if RCDM_ATTN_TRACE
constexpr int ATTN_TRACE_CTAS = 2560, ATTN_TRACE_TILES = 64;
__device__ long long g_attn_trace[(size_t)ATTN_TRACE_CTAS * ATTN_TRACE_TILES * 16];
__device__ int g_attn_smid[ATTN_TRACE_CTAS];
#define ATTN_STAMP(tile, slot)                                                                                      \
  do {                                                                                                              \
    const int cta_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                                \
    if (cta_ < ATTN_TRACE_CTAS && (tile) < ATTN_TRACE_TILES)                                                        \
      g_attn_trace[((size_t)cta_ * ATTN_TRACE_TILES + (tile)) * 16 + (slot)] = clock64();                           \
  } while (0)
#else
#define ATTN_STAMP(tile, slot) do { } while (0)
#endif

template <int DPAD> struct Attn4Cfg {
  static constexpr int BLOCK_M = 128, BLOCK_N = 64;
  static constexpr int NCH = DPAD / 8;
  static constexpr int Q_BYTES = NCH * BLOCK_M * 16;
  static constexpr int KV_BYTES = NCH * BLOCK_N * 16;
  static constexpr int KV_STAGES = 2;
  static constexpr int P_BYTES = (BLOCK_N / 8) * BLOCK_M * 16;
  static constexpr int TMEM_COLS = (BLOCK_N + DPAD) <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_STAGES * KV_BYTES + 2 * P_BYTES + 1024 + 256;
  static constexpr int CTAS_PER_SM = DPAD <= 48 ? 3 : (DPAD <= 80 ? 2 : 1);
  static constexpr int THREADS = 192;  // 4 softmax warps, MMA issuer, loader
};

// MSUM (compile time): the head dim leaves a spare padded column (d < DPAD), so the P V MMA accumulates the row sums
// and the softmax threads carry no fp32 sums at all.
template <typename T, int DPAD, bool MSUM>
__global__ void __launch_bounds__(Attn4Cfg<DPAD>::THREADS, Attn4Cfg<DPAD>::CTAS_PER_SM)
flash_attn4_kernel(const __grid_constant__ AttnMaps maps, const AttnParams p) {
  using Cfg = Attn4Cfg<DPAD>;
  constexpr int KV = Cfg::KV_STAGES;
  constexpr int BN = Cfg::BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + KV * Cfg::KV_BYTES;
  uint8_t* sP = sV + KV * Cfg::KV_BYTES;  // [2][P_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * Cfg::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;           // [KV] K tile landed
  uint64_t* v_full = bars + 1 + KV;      // [KV] V tile landed
  uint64_t* k_empty = bars + 1 + 2 * KV; // [KV] Q K_j^T has read the K stage
  uint64_t* v_empty = bars + 1 + 3 * KV; // [KV] P_j V_j has read the V stage (also: O is quiescent up to tile j)
  uint64_t* s_full = bars + 1 + 4 * KV;  // S_j is in TMEM
  uint64_t* s_free = s_full + 1;         // every softmax thread holds S_j in registers
  uint64_t* p_full = s_full + 2;         // [2]
  uint64_t* o_done = s_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int n_kv = (p.S_kv + BN - 1) / BN;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < KV; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      mbar_init(s_full, 1);
      // one arrival per softmax WARP (after a warp sync), not per thread: 128 arrivals serialise on the barrier word
      mbar_init(s_free, 4);
      mbar_init(&p_full[0], 4);
      mbar_init(&p_full[1], 4);
      mbar_init(o_done, 1);
      fence_mbar_init();
      tma_prefetch_desc(&maps.q);
      tma_prefetch_desc(&maps.k);
      tma_prefetch_desc(&maps.v);
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + BN;
  pdl_sync();

  if (warp == 5) {
    // =================================== loader warp (one elected lane) ===================================
    // A thread that issues a cp.async.bulk.tensor is held ~500 cycles by it (timeline trace, scripts/attn_trace.py: with the
    // loads issued by the MMA thread, that thread's serial work per K/V tile - Q K^T issue 350 + two loads 1080 + P V issue
    // 480 cycles - WAS the tile period, 2400 cycles, and the softmax warps idled 35 % of the time waiting for S), so the
    // loads have a warp of their own.
    if (elect_one()) {
      // V box: the real chunks only when the P V MMA sums the rows through the (pre-set) padded chunk
      const uint32_t v_bytes = MSUM ? (uint32_t)(p.d / 8) * (BN * 16) : (uint32_t)Cfg::KV_BYTES;
      auto load_k = [&](int t) {
        const int s = t % KV;
#if RCDM_ATTN_EXPERIMENT == 5
        if (t >= KV) {  // bottleneck analysis only: no K/V traffic after the first tiles
          mbar_expect_tx(&k_full[s], 0);
          return;
        }
#endif
        mbar_expect_tx(&k_full[s], Cfg::KV_BYTES);
        tma_load_5d(sK + s * Cfg::KV_BYTES, &maps.k, &k_full[s], 0, t * BN, 0, head, img);
      };
      auto load_v = [&](int t) {
        const int s = t % KV;
#if RCDM_ATTN_EXPERIMENT == 5
        if (t >= KV) {
          mbar_expect_tx(&v_full[s], 0);
          return;
        }
#endif
        mbar_expect_tx(&v_full[s], v_bytes);
        tma_load_5d(sV + s * Cfg::KV_BYTES, &maps.v, &v_full[s], 0, t * BN, 0, head, img);
      };
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      tma_load_5d(sQ, &maps.q, q_full, 0, q_tile * 128, 0, head, img);
      for (int t = 0; t < KV && t < n_kv; ++t) load_k(t);
      for (int t = 0; t < KV && t < n_kv; ++t) load_v(t);
      // refills in the order the stages come free: K_{j+KV} into the stage Q K_j^T released (it completed before S_j was
      // handed to the softmax threads), V_{j-1+KV} into the stage P_{j-1} V_{j-1} released (issued at the end of tile j-1)
      for (int j = 0; j < n_kv; ++j) {
        if (j + KV < n_kv) {
          mbar_wait(&k_empty[j % KV], (j / KV) & 1);
          load_k(j + KV);
        }
        if (j >= 1 && j - 1 + KV < n_kv) {
          mbar_wait(&v_empty[(j - 1) % KV], ((j - 1) / KV) & 1);
          load_v(j - 1 + KV);
        }
      }
    }
  } else if (warp == 4) {
    // =================================== MMA issuer (one elected lane) ===================================
    // This thread shares its scheduler with softmax warp 0 of every resident CTA (timeline trace: warp 0 lagged the other
    // three by ~300 cycles per tile while this loop rebuilt its 14 shared-memory descriptors per tile, ~220 instructions),
    // so everything that does not change is computed once: the descriptors of Q, of both K / V stages and of both P
    // buffers live in registers, and the tile loop is unrolled by the stage / buffer parity.
    if (elect_one()) {
      static_assert(KV == 2, "the MMA loop is unrolled by the parity of the two K/V stages and P buffers");
      constexpr uint32_t idesc_qk = umma_idesc_f16(DT<T>::umma_fmt, 128, BN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(DT<T>::umma_fmt, 128, DPAD, 0, 1);  // B (=V) is MN-major
      constexpr int NQK = DPAD / 16, NPV = BN / 16;
      uint64_t qd[NQK], kd[2][NQK], pd[2][NPV], vd[2][NPV];
#pragma unroll
      for (int ks = 0; ks < NQK; ++ks) {
        qd[ks] = umma_smem_desc(smem_u32(sQ) + ks * 2 * (128 * 16), 128 * 16, 128, UMMA_SWIZZLE_NONE);
#pragma unroll
        for (int st = 0; st < 2; ++st)
          kd[st][ks] = umma_smem_desc(smem_u32(sK + st * Cfg::KV_BYTES) + ks * 2 * (BN * 16), BN * 16, 128, UMMA_SWIZZLE_NONE);
      }
#pragma unroll
      for (int ks = 0; ks < NPV; ++ks)
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          pd[st][ks] = umma_smem_desc(smem_u32(sP + st * Cfg::P_BYTES) + ks * 4096, 2048, 128, UMMA_SWIZZLE_NONE);
          vd[st][ks] = umma_smem_desc(smem_u32(sV + st * Cfg::KV_BYTES) + ks * 256, 128, BN * 16, UMMA_SWIZZLE_NONE);
        }
      // S_t = Q K_t^T into the (single) S buffer; st = t % 2 at compile time
      auto mma_qk = [&](int t, auto st_c) {
        constexpr int st = decltype(st_c)::value;
        mbar_wait(&k_full[st], (t >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < NQK; ++ks) umma_f16_ss(tmem_S, qd[ks], kd[st][ks], idesc_qk, ks != 0);
        umma_commit(&k_empty[st]);  // the K stage is free once these MMAs have read it
        umma_commit(s_full);
      };
      // tile j with j % 2 == par: hand S_{j+1} to the tensor core as soon as S_j is in registers, then O += P_j V_j
      auto step = [&](int j, auto par_c) {
        constexpr int par = decltype(par_c)::value;
        if (j + 1 < n_kv) {
          mbar_wait(s_free, j & 1);  // S_j now lives in the softmax threads' registers: the TMEM buffer is free
          ATTN_STAMP(j, 12);
          tc_fence_after();
          mma_qk(j + 1, std::integral_constant<int, 1 - par>{});  // overlaps the exponentials of tile j
          ATTN_STAMP(j, 13);
        }
        // the one long wait of this thread (the softmax of tile j): sleep between polls, the scheduler's issue slots
        // belong to softmax warp 0
        mbar_wait_backoff(&p_full[par], (j >> 1) & 1, 64);
        ATTN_STAMP(j, 14);
        mbar_wait(&v_full[par], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < NPV; ++ks) umma_f16_ss(tmem_O, pd[par][ks], vd[par][ks], idesc_pv, (j | ks) != 0);
        umma_commit(&v_empty[par]);  // P_j V_j done: the V stage is free (also what the rare O rescale waits for)
        ATTN_STAMP(j, 15);
        if (j == n_kv - 1) umma_commit(o_done);
      };
      mbar_wait(q_full, 0);
      mma_qk(0, std::integral_constant<int, 0>{});
      for (int j = 0; j < n_kv; j += 2) {
        step(j, std::integral_constant<int, 0>{});
        if (j + 1 < n_kv) step(j + 1, std::integral_constant<int, 1>{});
      }
    }
  } else {
    // =================================== softmax warps: one thread per query row ===================================
    const int row = warp * 32 + lane;
    const uint32_t lane_sel = uint32_t(warp * 32) << 16;
    const float sc = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    constexpr bool mma_sum = MSUM;  // spare padded column of V carries ones: the P V MMA accumulates the row sums
    using T2 = typename DT<T>::T2;
    if (mma_sum && row < BN) {
      // the padded chunk of every V stage (never touched by TMA: the V box ends at the last real chunk): [1, 0, ..., 0];
      // made visible to the tensor core by the fence that precedes this thread's first p_full arrival
      uint4 one = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<T*>(&one) = DT<T>::from_f(1.0f);
      for (int c = p.d / 8; c < DPAD / 8; ++c)  // (further padded chunks, if any: zeros)
#pragma unroll
        for (int s = 0; s < KV; ++s)
          *reinterpret_cast<uint4*>(sV + s * Cfg::KV_BYTES + c * (BN * 16) + row * 16) =
              c == p.d / 8 ? one : make_uint4(0u, 0u, 0u, 0u);
    }

#if RCDM_ATTN_TRACE
    if (threadIdx.x == 0) {
      unsigned smid_;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
      const int cta_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      if (cta_ < ATTN_TRACE_CTAS) g_attn_smid[cta_] = (int)smid_;
    }
#endif
    for (int j = 0; j < n_kv; ++j) {
      if (lane == 0) ATTN_STAMP(j, warp);
      mbar_wait(s_full, j & 1);
      if (lane == 0) ATTN_STAMP(j, 4 + warp);
      tc_fence_after();
      uint32_t r[BN];
      tmem_ld32(tmem_S + lane_sel, r);
      tmem_ld32(tmem_S + lane_sel + 32, r + 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      uint8_t* sP_row = sP + (j & 1) * Cfg::P_BYTES + row * 16;
      const int kv_valid = min(BN, p.S_kv - j * BN);
      if (kv_valid < BN) {
#pragma unroll
        for (int i = 0; i < BN; ++i)
          if (i >= kv_valid) r[i] = 0xff800000u;  // -inf: exp2 -> 0
      }
      auto row_max = [&]() -> float {
        float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < BN; i += 2) {
          x0 = fmaxf(x0, __uint_as_float(r[i]));
          x1 = fmaxf(x1, __uint_as_float(r[i + 1]));
        }
        return fmaxf(x0, x1);
      };
      // exponentiate the row against `mref`, write P; returns the row sum (0 when the MMA sums) and the largest
      // probability written (in pmax)
      auto exp_pass = [&](float mref, float& pmax) -> float {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        T2 mx2 = DT<T>::from_f2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < BN / 8; ++g) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x = fmaf(__uint_as_float(r[g * 8 + i]), sc, -mref);
            // interleave the polynomial elements between the MUFU ones (1, 4, 6, 3 of each 8)
            constexpr int order[8] = {1, 4, 6, 3, 0, 2, 5, 7};
            bool poly = false;
#pragma unroll
            for (int q = 0; q < RCDM_ATTN_POLY_PER8; ++q) poly |= (order[q] == i);
            pv[i] = poly ? exp2_poly(x) : exp2f(x);
          }
          if constexpr (!mma_sum) {
            s0 += pv[0] + pv[4];
            s1 += pv[1] + pv[5];
            s2 += pv[2] + pv[6];
            s3 += pv[3] + pv[7];
          }
          uint4 pk = pack8<T>(pv);
          const T2* p2 = reinterpret_cast<const T2*>(&pk);
          mx2 = __hmax2(mx2, __hmax2(__hmax2(p2[0], p2[1]), __hmax2(p2[2], p2[3])));
          *reinterpret_cast<uint4*>(sP_row + g * 2048) = pk;  // P tile, K-major interleaved: [kv/8][row][8]
        }
        const float2 mxf = DT<T>::to_f2(mx2);
        pmax = fmaxf(mxf.x, mxf.y);
        return (s0 + s1) + (s2 + s3);
      };
      float pmax;
      if (j == 0) {
        m_run = row_max() * sc;
        l_run = exp_pass(m_run, pmax);
      } else {
        const float sum = exp_pass(m_run, pmax);
        if (__any_sync(0xffffffffu, pmax > 256.0f)) {
          // rare: some probability left the comfortable 16-bit range -> re-reference to the new maximum
          const float m_new = fmaxf(m_run, row_max() * sc);
          const float alpha = exp2f(m_run - m_new);
          const float sum2 = exp_pass(m_new, pmax);
          // O (incl. the row-sum column) must be quiescent: P_{j-1} V_{j-1} done = its K/V stage released.  (That
          // barrier cannot run ahead: its next completion needs P_{j-1+KV} from these same threads.)
          mbar_wait(&v_empty[(j - 1) % KV], ((j - 1) / KV) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < DPAD; c += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_O + lane_sel + c, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_O + lane_sel + c, o);
          }
          tmem_wait_st();
          l_run = l_run * alpha + sum2;
          m_run = m_new;
        } else {
          l_run += sum;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ATTN_STAMP(j, 8 + warp);
        mbar_arrive(&p_full[j & 1]);
      }
    }
    // ---- normalise and store
    mbar_wait(o_done, 0);
    tc_fence_after();
    if (mma_sum) {
      uint32_t o[16];
      tmem_ld16(tmem_O + lane_sel + (p.d & ~15), o);
      tmem_wait_ld();
      l_run = __uint_as_float(o[p.d & 15]);
    }
    const float inv_l = 1.0f / l_run;
    const int qrow = q_tile * 128 + row;
    T* out = reinterpret_cast<T*>(p.out) + ((size_t)img * p.S_q + qrow) * p.ldo + head * p.d;
#pragma unroll 1
    for (int c = 0; c < DPAD; c += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_O + lane_sel + c, o);
      tmem_wait_ld();
      if (qrow < p.S_q) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(o[i]) * inv_l;
        if (c < p.d) *reinterpret_cast<uint4*>(out + c) = pack8<T>(v);
        if (c + 8 < p.d) *reinterpret_cast<uint4*>(out + c + 8) = pack8<T>(v + 8);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// temporal attention: qkv [rows, 3C] with rows ordered (b, f, hw); out [rows, C]
// one thread per (b, hw, head); F <= 8 frames
// ------------------------------------------------------------------------------------------
template <typename T, int F>
__global__ void temporal_attn_kernel(const T* __restrict__ qkv, T* __restrict__ out, int batch, int hw, int heads,
                                     int d, float scale) {
  const int C = heads * d;
  const size_t total = (size_t)batch * hw * heads;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_sync();
  if (idx >= total) return;
  const int h = (int)(idx % heads);
  const size_t loc = idx / heads;
  const int pix = (int)(loc % hw);
  const int b = (int)(loc / hw);
  const size_t row0 = (size_t)b * F * hw + pix;  // frame f -> row0 + f*hw
  const int ld = 3 * C;
  float s[F][F];
#pragma unroll
  for (int i = 0; i < F; ++i)
#pragma unroll
    for (int j = 0; j < F; ++j) s[i][j] = 0.f;
  for (int c = 0; c < d; c += 8) {
    float qf[F][8], kf[F][8];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const T* base = qkv + (row0 + (size_t)f * hw) * ld + h * d + c;
      unpack8<T>(__ldg(reinterpret_cast<const uint4*>(base)), qf[f]);
      unpack8<T>(__ldg(reinterpret_cast<const uint4*>(base + C)), kf[f]);
    }
#pragma unroll
    for (int i = 0; i < F; ++i)
#pragma unroll
      for (int j = 0; j < F; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) s[i][j] += qf[i][e] * kf[j][e];
  }
#pragma unroll
  for (int i = 0; i < F; ++i) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      s[i][j] *= scale;
      mx = fmaxf(mx, s[i][j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      s[i][j] = __expf(s[i][j] - mx);
      sum += s[i][j];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < F; ++j) s[i][j] *= inv;
  }
  for (int c = 0; c < d; c += 8) {
    float vf[F][8];
#pragma unroll
    for (int f = 0; f < F; ++f)
      unpack8<T>(__ldg(reinterpret_cast<const uint4*>(qkv + (row0 + (size_t)f * hw) * ld + 2 * C + h * d + c)), vf[f]);
#pragma unroll
    for (int i = 0; i < F; ++i) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < F; ++j) acc += s[i][j] * vf[j][e];
        o[e] = acc;
      }
      *reinterpret_cast<uint4*>(out + (row0 + (size_t)i * hw) * C + h * d + c) = pack8<T>(o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// temporal attention, tiled: one CTA = PT consecutive pixels of one clip half b.  The F frame blocks
// [PT rows x 3C] of the fused QKV matrix are contiguous in memory -> copied to shared memory with coalesced 16-byte
// cp.async; thread (pixel, head, query frame i) computes one row of the F x F attention from shared memory and
// overwrites its own q_i slice with o_i (no other thread reads q_i); the q-third of the tile is then written back
// with coalesced 16-byte stores.  HBM/L2 traffic = the algorithmic minimum (read 3C, write C per row).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// DH = compile-time head dim (40 / 80 / 160: loops fully unrolled, address arithmetic folded; the kernel is bound by
// instruction issue, ncu: 70 % issue-active) or 0 = runtime `d_rt`.
template <typename T, int F, int DH>
__global__ void __launch_bounds__(512)
temporal_attn_tile_kernel(const T* __restrict__ qkv, T* __restrict__ out, int hw, int heads, int d_rt, int PT,
                          float scale) {
  extern __shared__ __align__(16) uint8_t tsm[];
  T* sm = reinterpret_cast<T*>(tsm);  // [F][PT][3C]
  const int d = DH > 0 ? DH : d_rt;
  const int C = heads * d, ld = 3 * C;
  const int tiles_per_b = hw / PT;
  const int b = blockIdx.x / tiles_per_b, pix0 = (blockIdx.x - b * tiles_per_b) * PT;
  const size_t row_b = (size_t)b * F * hw + pix0;
  const int chunks_per_f = PT * ld / 8;
  pdl_sync();
  for (int idx = threadIdx.x; idx < F * chunks_per_f; idx += blockDim.x) {
    const int f = idx / chunks_per_f, c = idx - f * chunks_per_f;
    cp_async16(sm + (size_t)f * PT * ld + (size_t)c * 8, qkv + (row_b + (size_t)f * hw) * ld + (size_t)c * 8);
  }
  cp_async_wait_all();
  __syncthreads();
  if (threadIdx.x < PT * heads * F) {
    const int i = threadIdx.x % F;
    const int h = (threadIdx.x / F) % heads;
    const int p = threadIdx.x / (F * heads);
    T* qp = sm + (size_t)(i * PT + p) * ld + h * d;
    float s[F];
#pragma unroll
    for (int j = 0; j < F; ++j) s[j] = 0.f;
    const T* kbase = sm + (size_t)p * ld + C + h * d;       // + j * PT * ld per frame; V at + C
    const int fstride = PT * ld;
#pragma unroll
    for (int c = 0; c < (DH > 0 ? DH : 1 << 30); c += 8) {
      if (DH == 0 && c >= d) break;
      float qf[8];
      unpack8<T>(*reinterpret_cast<const uint4*>(qp + c), qf);
#pragma unroll
      for (int j = 0; j < F; ++j) {
        float kf[8];
        unpack8<T>(*reinterpret_cast<const uint4*>(kbase + j * fstride + c), kf);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[j] = fmaf(qf[e], kf[e], s[j]);
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      s[j] *= scale;
      mx = fmaxf(mx, s[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      s[j] = __expf(s[j] - mx);
      sum += s[j];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < F; ++j) s[j] *= inv;
#pragma unroll
    for (int c = 0; c < (DH > 0 ? DH : 1 << 30); c += 8) {
      if (DH == 0 && c >= d) break;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) {
        float vf[8];
        unpack8<T>(*reinterpret_cast<const uint4*>(kbase + C + j * fstride + c), vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(s[j], vf[e], o[e]);
      }
      *reinterpret_cast<uint4*>(qp + c) = pack8<T>(o);  // in place: only this thread ever reads q_i
    }
  }
  __syncthreads();
  const int cvec = C / 8;
  for (int idx = threadIdx.x; idx < F * PT * cvec; idx += blockDim.x) {
    const int c8 = idx % cvec;
    const int fp = idx / cvec;  // f * PT + p
    const int f = fp / PT, p = fp - f * PT;
    *reinterpret_cast<uint4*>(out + (row_b + (size_t)f * hw + p) * C + (size_t)c8 * 8) =
        *reinterpret_cast<const uint4*>(sm + (size_t)fp * ld + (size_t)c8 * 8);
  }
}

}  // namespace rcdm
