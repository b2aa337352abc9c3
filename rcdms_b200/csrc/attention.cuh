// Attention kernels for sm_100a.
//
// flash_attn_kernel: softmax(scale * Q K^T) V without materialising the scores (the reference writes an
// (80, 4096, 4096) score tensor per 64x64 layer: attention.py:170-199).  One CTA = 128 queries of one
// (image, head).  Both GEMMs run on tcgen05 with accumulators in TMEM:
//     S = Q K_j^T   (128 x 128 fp32, TMEM cols [0,128))      O += P_j V_j   (128 x DPAD fp32, TMEM cols [128, ...))
// Q/K/V head slices are fetched straight out of the fused projection output [rows, ld] by 5-D TMA maps
// (8 elems, row, 16-byte chunk, head, image) into the no-swizzle "interleaved" UMMA layout
// [chunk][row][8 elems]; out-of-range chunks / rows are zero-filled by TMA, which pads head_dim 40 -> 48 and
// ragged KV lengths (context L = 85 / 91) for free.  V is consumed as an MN-major B operand, so no transpose.
// Warps 0-3: online softmax (thread <-> query row), P written to smem as the A operand of the second GEMM.
// Warp 4 (one lane): TMA producer + MMA issuer.
//
// temporal_attn_kernel: the motion modules' attention over the f = 5 frames at each spatial location
// (motion_module.py:294-354): a 5x5 problem per (location, head) -> CUDA cores, one thread per (location, head),
// q/k/v read once directly from the (b f hw)-ordered token matrix (no "(b f) d c -> (b d) f c" copies).
#pragma once
#include <type_traits>

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

template <int DPAD> struct AttnCfg {
  static constexpr int NCH = DPAD / 8;               // 16-byte chunks per head row
  static constexpr int TILE_BYTES = NCH * 128 * 16;  // one 128-row operand tile
  static constexpr int KV_STAGES = DPAD <= 80 ? 2 : 1;
  static constexpr int P_BYTES = 16 * 128 * 16;
  static constexpr int TMEM_COLS = (128 + DPAD) <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * KV_STAGES) + P_BYTES + 1024 + 256;
};

template <typename T, int DPAD>
__global__ void __launch_bounds__(160, AttnCfg<DPAD>::SMEM_BYTES <= 113 * 1024 ? 2 : 1)
flash_attn_kernel(const __grid_constant__ AttnMaps maps, const AttnParams p) {
  using Cfg = AttnCfg<DPAD>;
  constexpr int KV = Cfg::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::TILE_BYTES;
  uint8_t* sV = sK + KV * Cfg::TILE_BYTES;
  uint8_t* sP = sV + KV * Cfg::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;        // [KV]
  uint64_t* kv_empty = bars + 1 + KV;  // [KV]
  uint64_t* s_full = bars + 1 + 2 * KV;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_done = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int n_kv = (p.S_kv + 127) / 128;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < KV; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(p_full, 128);
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
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 4) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(DT<T>::umma_fmt, 128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(DT<T>::umma_fmt, 128, DPAD, 0, 1);  // B (=V) is MN-major
      mbar_expect_tx(q_full, Cfg::TILE_BYTES);
      tma_load_5d(sQ, &maps.q, q_full, 0, q_tile * 128, 0, head, img);
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      for (int j = 0; j < n_kv; ++j) {
        // ---- producer: keep the K/V ring full
        const int t_first = (j == 0) ? 0 : j + KV - 1, t_last = j + KV - 1;
        for (int t = t_first; t <= t_last; ++t) {
          if (t >= n_kv) break;
          const int s = t % KV;
          if (t >= KV) mbar_wait(&kv_empty[s], ((t / KV) - 1) & 1);
          mbar_expect_tx(&kv_full[s], 2 * Cfg::TILE_BYTES);
          tma_load_5d(sK + s * Cfg::TILE_BYTES, &maps.k, &kv_full[s], 0, t * 128, 0, head, img);
          tma_load_5d(sV + s * Cfg::TILE_BYTES, &maps.v, &kv_full[s], 0, t * 128, 0, head, img);
        }
        const int s = j % KV;
        if (j == 0) mbar_wait(q_full, 0);
        mbar_wait(&kv_full[s], (j / KV) & 1);
        tc_fence_after();
        // ---- S = Q K^T : A, B K-major interleaved; +2 chunks (= 4096 B) per K=16 step
        const uint32_t k_addr = smem_u32(sK + s * Cfg::TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < DPAD / 16; ++ks) {
          const uint64_t ad = umma_smem_desc(q_addr + ks * 4096, 2048, 128, UMMA_SWIZZLE_NONE);
          const uint64_t bd = umma_smem_desc(k_addr + ks * 4096, 2048, 128, UMMA_SWIZZLE_NONE);
          umma_f16_ss(tmem_S, ad, bd, idesc_qk, ks != 0);
        }
        umma_commit(s_full);
        // ---- O += P V : A = P K-major interleaved, B = V MN-major interleaved (+16 kv rows = 256 B per step)
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + s * Cfg::TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = umma_smem_desc(p_addr + ks * 4096, 2048, 128, UMMA_SWIZZLE_NONE);
          const uint64_t bd = umma_smem_desc(v_addr + ks * 256, 128, 2048, UMMA_SWIZZLE_NONE);
          umma_f16_ss(tmem_O, ad, bd, idesc_pv, (j | ks) != 0);
        }
        umma_commit(&kv_empty[s]);
        if (j == n_kv - 1) umma_commit(o_done);
      }
    }
  } else {
    // =================================== softmax warps ===================================
    // One thread per query row.  Lazy online softmax: the first KV tile fixes the reference maximum m_run with a
    // separate max pass; every later tile is ONE pass that exponentiates against m_run while tracking the tile
    // maximum, and only if some row's tile maximum exceeds m_run by more than 2^8 (P would approach the 16-bit
    // range) does the warp fall back to re-exponentiating the tile and rescaling O in TMEM.
    const int row = warp * 32 + lane;
    const uint32_t lane_sel = uint32_t(warp * 32) << 16;
    const float sc = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* sP_row = sP + row * 16;
    // Row sums for free: when the head dim leaves a spare padded column (d = 40 in a 48-wide tile), column d of
    // every V row is set to 1 so that the P V MMA accumulates sum_j P_ij (of the ROUNDED probabilities) in TMEM.
    const bool mma_sum = p.d < DPAD;
    using T2 = typename DT<T>::T2;

    // exponentiate one 128-column S row against `mref`, write P; returns the row sum (0 when MS) and, in `pmax`,
    // the largest probability written (as float)
    auto exp_pass = [&](auto ms_tag, float mref, int kv_valid, float& pmax) -> float {
      constexpr bool MS = decltype(ms_tag)::value;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      T2 mx2 = DT<T>::from_f2(0.f, 0.f);
      const bool full = kv_valid >= 128;
      // software pipeline over four 32-column chunks: the TMEM load of chunk k+1 is in flight while chunk k is
      // exponentiated (tcgen05.wait::ld only before the data is consumed)
      uint32_t rbuf[2][32];
      tmem_ld32(tmem_S + lane_sel, rbuf[0]);
      tmem_wait_ld();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t* r = rbuf[k & 1];
        if (k < 3) tmem_ld32(tmem_S + lane_sel + (k + 1) * 32, rbuf[(k + 1) & 1]);
        const int c = k * 32;
        if (!full) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i >= kv_valid) r[i] = 0xff800000u;  // -inf: exp2 -> 0
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pv[i] = exp2f(fmaf(__uint_as_float(r[g * 8 + i]), sc, -mref));
          if constexpr (!MS) {
            s0 += pv[0] + pv[4];
            s1 += pv[1] + pv[5];
            s2 += pv[2] + pv[6];
            s3 += pv[3] + pv[7];
          }
          uint4 pk = pack8<T>(pv);
          const T2* p2 = reinterpret_cast<const T2*>(&pk);
          mx2 = __hmax2(mx2, __hmax2(__hmax2(p2[0], p2[1]), __hmax2(p2[2], p2[3])));
          // P tile, K-major interleaved: [chunk = kv/8][row][8 elems]
          *reinterpret_cast<uint4*>(sP_row + (c / 8 + g) * 2048) = pk;
        }
        if (k < 3) tmem_wait_ld();
      }
      const float2 mxf = DT<T>::to_f2(mx2);
      pmax = fmaxf(mxf.x, mxf.y);
      return (s0 + s1) + (s2 + s3);
    };
    auto row_max = [&](int kv_valid) -> float {
      float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 128; c += 64) {
        uint32_t r[64];
        tmem_ld32(tmem_S + lane_sel + c, r);
        tmem_ld32(tmem_S + lane_sel + c + 32, r + 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          if (c + i < kv_valid) x0 = fmaxf(x0, __uint_as_float(r[i]));
          if (c + i + 1 < kv_valid) x1 = fmaxf(x1, __uint_as_float(r[i + 1]));
        }
      }
      return fmaxf(x0, x1);
    };
    auto run_pass = [&](float mref, int kv_valid, float& pmax) -> float {
      return mma_sum ? exp_pass(std::true_type{}, mref, kv_valid, pmax) : exp_pass(std::false_type{}, mref, kv_valid, pmax);
    };

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv_valid = min(128, p.S_kv - j * 128);
      if (mma_sum)  // V tile j has landed (same barrier as K): set its "ones" column for this thread's kv row
        *reinterpret_cast<T*>(sV + (j % KV) * Cfg::TILE_BYTES + (p.d / 8) * 2048 + row * 16) = DT<T>::from_f(1.0f);
      float pmax;
      if (j == 0) {
        m_run = row_max(kv_valid) * sc;
        l_run = run_pass(m_run, kv_valid, pmax);
      } else {
        const float sum = run_pass(m_run, kv_valid, pmax);
        if (__any_sync(0xffffffffu, pmax > 256.0f)) {
          // rare: some probability left the comfortable 16-bit range -> re-reference to the new maximum
          const float m_new = fmaxf(m_run, row_max(kv_valid) * sc);
          const float alpha = exp2f(m_run - m_new);
          const float sum2 = run_pass(m_new, kv_valid, pmax);
#pragma unroll 1
          for (int c = 0; c < DPAD; c += 16) {  // rescales the row-sum column as well
            uint32_t r[16];
            tmem_ld16(tmem_O + lane_sel + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(tmem_O + lane_sel + c, r);
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
      mbar_arrive(p_full);
    }
    // ---- normalise and store
    mbar_wait(o_done, 0);
    tc_fence_after();
    if (mma_sum) {
      uint32_t r[16];
      tmem_ld16(tmem_O + lane_sel + (p.d & ~15), r);
      tmem_wait_ld();
      l_run = __uint_as_float(r[p.d & 15]);
    }
    const float inv_l = 1.0f / l_run;
    const int qrow = q_tile * 128 + row;
    T* out = reinterpret_cast<T*>(p.out) + ((size_t)img * p.S_q + qrow) * p.ldo + head * p.d;
#pragma unroll 1
    for (int c = 0; c < DPAD; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_O + lane_sel + c, r);
      tmem_wait_ld();
      if (qrow < p.S_q) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * inv_l;
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

}  // namespace rcdm
