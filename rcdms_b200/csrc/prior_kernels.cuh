// Kernels that only the stage-1 frame-prior loop needs (reference:
// src/models/myprior_transformer.py:275-411, src/pipelines/prior_pipeline.py:283-352).  The heavy work of the prior
// (20 x [LayerNorm -> QKV -> attention -> out-proj -> GELU feed-forward] + 20 prior-state motion modules) runs on the
// same tcgen05 GEMM / LayerNorm / temporal-attention kernels as the UNet; what is new here is small and HBM / latency
// bound: the masked (causal + key padding) attention over <= 256 tokens, the per-step token assembly and the fused
// CFG + UnCLIP scheduler step.
#pragma once
#include "common.cuh"

namespace rcdm {

// ---------------------------------------------------------------------------------------------------------------
// Masked self-attention over a short sequence (CrossAttention._attention with an additive mask,
// attention.py:171-199; mask built at myprior_transformer.py:160-165,386-390):
//   P = softmax(scale * Q K^T + key_bias[b, j] + (causal && j > i ? -10000 : 0)),  O = P V
// qkv: [(b, s), ld] with q | k | v at columns 0 | C | 2C (the fused projection output), head h at [h*d, h*d + d).
// One CTA per (batch, head, 32-query chunk): K and V of the head live in shared memory (K rows padded by one word so
// that lanes reading different keys hit different banks), one warp per query row, lane <-> key for the scores and
// lane <-> channel pair for P V.  Probabilities are normalised in fp32 and rounded to the storage dtype before P V,
// like the reference's `attention_probs.to(value.dtype)`.
// S <= 32 * MASKED_ATTN_KPL keys; d even, d <= 256.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MASKED_ATTN_KPL = 8;      // keys per lane (S <= 256)
constexpr int MASKED_ATTN_QCHUNK = 32;  // query rows per CTA
constexpr int MASKED_ATTN_WARPS = 4;

inline size_t masked_attn_smem_bytes(int S, int d) {
  return (size_t)S * (d + 2) * 2 + (size_t)S * d * 2 + (size_t)MASKED_ATTN_WARPS * d * 4 +
         (size_t)MASKED_ATTN_WARPS * ((S + 31) / 32 * 32) * 4;
}

template <typename T>
__global__ void __launch_bounds__(MASKED_ATTN_WARPS * 32)
masked_attn_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ key_bias, int causal,
                   T* __restrict__ out, int ldo, int S, int heads, int d, float scale) {
  using T2 = typename DT<T>::T2;
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int C = heads * d;
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int q0 = blockIdx.y * MASKED_ATTN_QCHUNK;
  const int kst = d + 2;  // padded K row pitch (elements); (d/2 + 1) words is odd for d % 4 == 0
  const int Sp = (S + 31) / 32 * 32;
  T* Ks = reinterpret_cast<T*>(sm_raw);
  T* Vs = Ks + (size_t)S * kst;
  float* qs = reinterpret_cast<float*>(Vs + (size_t)S * d);
  float* ps = qs + MASKED_ATTN_WARPS * d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_sync();
  // causal: this chunk's queries see keys 0 .. q0 + QCHUNK - 1 only
  const int kmax = causal ? min(S, q0 + MASKED_ATTN_QCHUNK) : S;
  const T* base = qkv + (size_t)b * S * ld + h * d;
  const int d2 = d >> 1;
  for (int i = tid; i < kmax * d2; i += blockDim.x) {
    const int j = i / d2, c = (i % d2) * 2;
    const T2 kv = *reinterpret_cast<const T2*>(base + (size_t)j * ld + C + c);
    const T2 vv = *reinterpret_cast<const T2*>(base + (size_t)j * ld + 2 * C + c);
    *reinterpret_cast<T2*>(Ks + (size_t)j * kst + c) = kv;
    *reinterpret_cast<T2*>(Vs + (size_t)j * d + c) = vv;
  }
  __syncthreads();
  float* qw = qs + warp * d;
  float* pw = ps + warp * Sp;
  for (int r = warp; r < MASKED_ATTN_QCHUNK; r += MASKED_ATTN_WARPS) {
    const int i = q0 + r;
    if (i >= S) break;  // warp-uniform
    for (int c = lane; c < d2; c += 32) {
      const float2 f = DT<T>::to_f2(*reinterpret_cast<const T2*>(base + (size_t)i * ld + 2 * c));
      qw[2 * c] = f.x * scale;
      qw[2 * c + 1] = f.y * scale;
    }
    __syncwarp();
    const int jend = causal ? min(kmax, i + 1) : kmax;  // keys above the diagonal carry -10000: exp underflows to 0
    float sc[MASKED_ATTN_KPL];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < MASKED_ATTN_KPL; ++t) {
      const int j = lane + 32 * t;
      sc[t] = -INFINITY;
      if (j < jend) {
        const T* kr = Ks + (size_t)j * kst;
        float a0 = 0.f, a1 = 0.f;
        for (int c = 0; c < d2; ++c) {
          const float2 kf = DT<T>::to_f2(*reinterpret_cast<const T2*>(kr + 2 * c));
          a0 = fmaf(qw[2 * c], kf.x, a0);
          a1 = fmaf(qw[2 * c + 1], kf.y, a1);
        }
        sc[t] = a0 + a1 + (key_bias ? __ldg(key_bias + (size_t)b * S + j) : 0.f);
        mx = fmaxf(mx, sc[t]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float l = 0.f;
#pragma unroll
    for (int t = 0; t < MASKED_ATTN_KPL; ++t) {
      const int j = lane + 32 * t;
      if (j < jend) {
        sc[t] = __expf(sc[t] - mx);
        l += sc[t];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float inv = 1.0f / l;
#pragma unroll
    for (int t = 0; t < MASKED_ATTN_KPL; ++t) {
      const int j = lane + 32 * t;
      if (j < jend) pw[j] = DT<T>::to_f(DT<T>::from_f(sc[t] * inv));
    }
    __syncwarp();
    for (int c = lane; c < d2; c += 32) {
      float o0 = 0.f, o1 = 0.f;
      for (int j = 0; j < jend; ++j) {
        const float2 vf = DT<T>::to_f2(*reinterpret_cast<const T2*>(Vs + (size_t)j * d + 2 * c));
        const float pj = pw[j];
        o0 = fmaf(pj, vf.x, o0);
        o1 = fmaf(pj, vf.y, o1);
      }
      *reinterpret_cast<T2*>(out + ((size_t)b * S + i) * ldo + h * d + 2 * c) = DT<T>::from_f2(o0, o1);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core version of the same attention for the prior's shape (head dim 64, <= 112 tokens): one warp owns 16
// query rows and the WHOLE key range, so scores, mask, softmax and P V all stay in registers (no online rescaling):
//   S = Q K^T   mma.sync.m16n8k16 (A = Q from global, B = K rows from shared memory, pitch d + 8 -> conflict-free)
//   P = round16(exp(S - max) / sum)   in the accumulator registers, which ARE the A fragments of the second MMA
//   O = P V     B = V^T from shared memory (transposed on the way in, pitch NT*8 + 8)
// One CTA per (batch, head); 4 warps, warp w takes the 16-row tiles w and (tiles - 1 - w), which balances the causal work.  Replaces ~50 M CUDA-core warp instructions of the generic
// kernel (70 us per launch, 13 % of a prior step) by ~25 k MMAs.
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct MmaOp;
template <> struct MmaOp<__half> {
  __device__ static __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct MmaOp<__nv_bfloat16> {
  __device__ static __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};

constexpr int MATTN_D = 64;    // head dim of this instantiation
constexpr int MATTN_NT = 14;   // 8-key tiles held in registers: S <= 112
constexpr int MATTN_KP = MATTN_D + 8;        // K row pitch (elements)
constexpr int MATTN_VP = MATTN_NT * 8 + 8;   // V^T row pitch (elements)
inline size_t masked_attn_mma_smem_bytes() {
  return (size_t)(MATTN_NT * 8) * MATTN_KP * 2 + (size_t)MATTN_D * MATTN_VP * 2;
}

constexpr int MATTN_THREADS = 128;  // 4 warps: warp w owns row tiles w and (tiles - 1 - w) -> equal causal work
template <typename T>
__global__ void __launch_bounds__(MATTN_THREADS, 4)
masked_attn_mma_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ key_bias, int causal,
                       T* __restrict__ out, int ldo, int S, int heads, float scale) {
  using T2 = typename DT<T>::T2;
  constexpr int D = MATTN_D, NT = MATTN_NT, KP = MATTN_KP, VP = MATTN_VP, SK = NT * 8;
  extern __shared__ __align__(16) uint8_t sm_raw[];
  T* Ks = reinterpret_cast<T*>(sm_raw);   // [SK][KP]   rows >= S zero
  T* Vt = Ks + SK * KP;                   // [D][VP]    columns >= S zero
  const int C = heads * D;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const T* base = qkv + (size_t)b * S * ld + h * D;
  pdl_sync();
  // causal: the CTA's last row tile sees keys < 16 * tiles, and row tile w only key tiles nt < 2 w + 2
  const int row_tiles = (S + 15) / 16;
  const int sk_used = causal ? min(SK, row_tiles * 16) : SK;
  const bool vec16 = (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(qkv) & 15) == 0);
  for (int i = tid; i < sk_used * (D / 8); i += blockDim.x) {  // 16-byte global loads: 8 channels of one key
    const int j = i / (D / 8), c = (i % (D / 8)) * 8;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (j < S) {
      const T* kp = base + (size_t)j * ld + C + c;
      if (vec16) {
        kv = __ldg(reinterpret_cast<const uint4*>(kp));
        vv = __ldg(reinterpret_cast<const uint4*>(kp + C));
      } else {
        uint32_t* k4 = reinterpret_cast<uint32_t*>(&kv);
        uint32_t* v4 = reinterpret_cast<uint32_t*>(&vv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          k4[e] = *reinterpret_cast<const uint32_t*>(kp + 2 * e);
          v4[e] = *reinterpret_cast<const uint32_t*>(kp + C + 2 * e);
        }
      }
    }
    *reinterpret_cast<uint4*>(Ks + j * KP + c) = kv;  // KP * 2 = 144 bytes: 16-byte aligned rows
    const T* ve = reinterpret_cast<const T*>(&vv);
#pragma unroll
    for (int e = 0; e < 8; ++e) Vt[(c + e) * VP + j] = ve[e];
  }
  __syncthreads();
  // no barrier below this point; every branch on `tile` is warp-uniform
  if (2 * warp > row_tiles - 1) return;  // warps beyond the middle tile have nothing to do
  for (int pass = 0; pass < 2; ++pass) {
  const int tile = pass == 0 ? warp : row_tiles - 1 - warp;
  if (pass == 1 && tile <= warp) break;  // the middle tile is its own partner
  const int r0 = tile * 16;
  const int nt_used = causal ? min(NT, 2 * tile + 2) : NT;
  const int row_a = r0 + g, row_b = r0 + g + 8;
  // ---- Q fragments (A operand, row-major 16 x 16 per k-step): a0 (g, 2t) a1 (g+8, 2t) a2 (g, 2t+8) a3 (g+8, 2t+8)
  uint32_t qa[D / 16][4];
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    const int c = ks * 16 + 2 * t;
    qa[ks][0] = row_a < S ? *reinterpret_cast<const uint32_t*>(base + (size_t)row_a * ld + c) : 0u;
    qa[ks][1] = row_b < S ? *reinterpret_cast<const uint32_t*>(base + (size_t)row_b * ld + c) : 0u;
    qa[ks][2] = row_a < S ? *reinterpret_cast<const uint32_t*>(base + (size_t)row_a * ld + c + 8) : 0u;
    qa[ks][3] = row_b < S ? *reinterpret_cast<const uint32_t*>(base + (size_t)row_b * ld + c + 8) : 0u;
  }
  // ---- S = Q K^T: accumulator tile nt holds (row g | g+8, keys nt*8 + 2t, +1)
  float sc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
    if (nt < nt_used) {
      const T* kr = Ks + (nt * 8 + g) * KP + 2 * t;  // B fragment: b0 (k = 2t, 2t+1; n = g), b1 (k + 8)
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks)
        MmaOp<T>::mma(sc[nt], qa[ks], *reinterpret_cast<const uint32_t*>(kr + ks * 16),
                      *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8));
    }
  }
  // ---- scale, additive mask, softmax over the full row (quad reduction: lanes 4g .. 4g+3 share rows g, g+8)
  float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + 2 * t + e;
      float kb = -INFINITY;  // padding keys, and (causal) key tiles entirely above this row tile's diagonal
      if (j < S && nt < nt_used) kb = key_bias ? __ldg(key_bias + (size_t)b * S + j) : 0.f;
      const float va = sc[nt][e] * scale + kb + ((causal && j > row_a) ? -10000.f : 0.f);
      const float vb = sc[nt][2 + e] * scale + kb + ((causal && j > row_b) ? -10000.f : 0.f);
      sc[nt][e] = va;
      sc[nt][2 + e] = vb;
      mxa = fmaxf(mxa, va);
      mxb = fmaxf(mxb, vb);
    }
  }
  mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 1));
  mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 2));
  mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 1));
  mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 2));
  float la = 0.f, lb = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      sc[nt][e] = __expf(sc[nt][e] - mxa);  // exp(-inf) = 0 for the padding keys
      sc[nt][2 + e] = __expf(sc[nt][2 + e] - mxb);
      la += sc[nt][e];
      lb += sc[nt][2 + e];
    }
  }
  la += __shfl_xor_sync(0xffffffffu, la, 1);
  la += __shfl_xor_sync(0xffffffffu, la, 2);
  lb += __shfl_xor_sync(0xffffffffu, lb, 1);
  lb += __shfl_xor_sync(0xffffffffu, lb, 2);
  const float ia = 1.0f / la, ib = 1.0f / lb;
  // ---- O = P V: the normalised, 16-bit-rounded probabilities of key tiles (2j, 2j+1) are the A fragment of k-step j
  float oc[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) oc[n][0] = oc[n][1] = oc[n][2] = oc[n][3] = 0.f;
#pragma unroll
  for (int kj = 0; kj < NT / 2; ++kj) {
    if (2 * kj >= nt_used) break;  // warp-uniform: the remaining probabilities are exactly zero
    uint32_t pa[4];
    T2 p0 = DT<T>::from_f2(sc[2 * kj][0] * ia, sc[2 * kj][1] * ia);
    T2 p1 = DT<T>::from_f2(sc[2 * kj][2] * ib, sc[2 * kj][3] * ib);
    T2 p2 = DT<T>::from_f2(sc[2 * kj + 1][0] * ia, sc[2 * kj + 1][1] * ia);
    T2 p3 = DT<T>::from_f2(sc[2 * kj + 1][2] * ib, sc[2 * kj + 1][3] * ib);
    pa[0] = *reinterpret_cast<uint32_t*>(&p0);
    pa[1] = *reinterpret_cast<uint32_t*>(&p1);
    pa[2] = *reinterpret_cast<uint32_t*>(&p2);
    pa[3] = *reinterpret_cast<uint32_t*>(&p3);
#pragma unroll
    for (int n = 0; n < D / 8; ++n) {
      const T* vr = Vt + (n * 8 + g) * VP + kj * 16 + 2 * t;  // b0 (keys kj*16 + 2t, +1; dim n*8 + g), b1 (keys + 8)
      MmaOp<T>::mma(oc[n], pa, *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
    }
  }
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int c = h * D + n * 8 + 2 * t;
    if (row_a < S) *reinterpret_cast<T2*>(out + ((size_t)b * S + row_a) * ldo + c) = DT<T>::from_f2(oc[n][0], oc[n][1]);
    if (row_b < S) *reinterpret_cast<T2*>(out + ((size_t)b * S + row_b) * ldo + c) = DT<T>::from_f2(oc[n][2], oc[n][3]);
  }
  }  // pass
}

// ---------------------------------------------------------------------------------------------------------------
// Cross-attention to a SHORT key range (the UNet's attn2: 85 / 91 context tokens, attention.py:140-168,170-199): the same
// register-resident scheme as masked_attn_mma_kernel - one warp owns 16 query rows and the whole key range, scores,
// softmax and P V never leave registers - with queries and keys from different tensors, no mask, any head dim D that
// is a multiple of 8 (k-steps padded to 16 with zeros), and a loop over the query tiles so that the K rows / V^T of one
// (image, head) are staged in shared memory once per CTA.  The tcgen05 flash kernel spends a whole 128-row CTA
// (TMEM allocation, barriers, 5-D TMA head gathers) on two key tiles: 45 us for 4096 x 85 keys against ~10 us of HBM time.
//   q[(img * S_q + i) * ldq + h * D + c],  k / v[(img * S_kv + j) * ldkv + h * D + c],  out like q with ldo
// grid (images * heads, query chunks); CTA = 4 warps; the chunk's 16-row tiles go round-robin over the warps.
// DEFAULT for key ranges <= 112 (library option attn_short_kv = 1; 0 = flash kernel).  History (scripts/bench_xattn.py,
// profiles/r02_xattn_bench*.txt; 80 images x 4096 queries x 91 keys, d = 40; flash kernel 292 us):
//   v1  4-byte fragment loads / stores straight from / to global memory: 333 us (L1 wavefronts: 8-16 lines per instruction)
//   v2  Q / O tiles through a per-warp shared-memory tile, 16 bytes per lane to global memory: 315 us; ncu
//       (profiles/r02_ncu_cross_attn.txt): no pipe saturated, ~1 035 warp instructions per 16-row tile of which 72 are MMAs:
//       144 B-fragment LDS.32, 112 FMUL (scale, normalise), 130 ISETP / FSEL (key mask on every element), 48 WARPSYNC and
//       predicated-off MMAs from the run-time tile count, 14 key tiles processed for 12 live ones
//   v3  (this) tile count NT a template parameter (8 / 11 / 12 exact, 14 generic), key mask on the last tile only, scale
//       folded into the exponent FFMA, normalisation deferred to the 16 x D output, K / V^T columns permuted inside each
//       16-block so that a lane's (b0, b1) fragment pair is ONE 8-byte load: 217 us (x 1024 d = 80: 113 vs 137 us flash;
//       x 256 d = 160: 84 vs 88 us; one clip, 10 images x 4096: 40 vs 48 us).
// q, out: 16-byte aligned rows (ldq, ldo multiples of 8).
// ---------------------------------------------------------------------------------------------------------------
constexpr int XATTN_NT = 14;                        // generic instantiation: S_kv <= 112
constexpr int XATTN_SK = XATTN_NT * 8;
constexpr int XATTN_VP = 112;                       // V^T row pitch (elements): >= 16 ceil(NT / 2), = 48 mod 64 (see xattn_kp)
constexpr int XATTN_THREADS = 128;
__host__ __device__ constexpr int xattn_dp(int D) { return (D + 15) / 16 * 16; }
// K row pitch (elements): the 8-byte B-fragment loads of a half-warp (rows g = 0..3 or 4..7, 8 t bytes into the row) are
// conflict-free when the pitch is 16 or 48 mod 64 elements
__host__ __device__ constexpr int xattn_kp(int D) { return (xattn_dp(D) % 64 == 16 || xattn_dp(D) % 64 == 48) ? xattn_dp(D) : xattn_dp(D) + 16; }
__host__ __device__ constexpr int xattn_qp(int D) { return xattn_dp(D) + 8; }  // Q / O staging pitch: 4-byte fragment accesses
__host__ __device__ constexpr size_t xattn_smem_bytes(int D, int NT) {  // K rows + V^T + one 16-row Q / O staging tile per warp
  return (size_t)NT * 8 * xattn_kp(D) * 2 + (size_t)xattn_dp(D) * XATTN_VP * 2 +
         (size_t)(XATTN_THREADS / 32) * 16 * xattn_qp(D) * 2;
}
// position of column c (0..15) inside its 16-block: (2t, 2t+1, 2t+8, 2t+9) -> 4t .. 4t+3
__host__ __device__ constexpr int xattn_perm16(int c) { return 4 * ((c & 7) >> 1) + 2 * (c >> 3) + (c & 1); }

// EXACT: 8 (NT - 1) < S_kv <= 8 NT (only the last key tile can hold dead keys); else any S_kv <= 8 NT.
template <typename T, int D, int NT, bool EXACT>
__global__ void __launch_bounds__(XATTN_THREADS)
cross_attn_mma_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, const T* __restrict__ v, int ldkv,
                      T* __restrict__ out, int ldo, int S_q, int S_kv, int heads, int tiles_per_cta, float scale_log2) {
  using T2 = typename DT<T>::T2;
  constexpr int DP = xattn_dp(D), KP = xattn_kp(D), QP = xattn_qp(D), VP = XATTN_VP;
  constexpr int NKJ = (NT + 1) / 2, NKP = NKJ * 16;  // 16-key steps of P V, keys they read
  constexpr int NCH = D / 8;                        // 8-wide output column tiles
  constexpr int NCHUNK = NCH > 10 ? 10 : NCH;       // P V in chunks of <= 80 output dims (register budget at D = 160)
  static_assert(NKP <= VP, "V^T pitch");
  extern __shared__ __align__(16) uint8_t sm_raw[];
  T* Ks = reinterpret_cast<T*>(sm_raw);   // [NT * 8][KP]  16-blocks permuted (xattn_perm16); rows >= S_kv, channels >= D zero
  T* Vt = Ks + NT * 8 * KP;               // [DP][VP]      key 16-blocks permuted; keys >= S_kv zero
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const T* kb = k + (size_t)b * S_kv * ldkv + h * D;
  const T* vb = v + (size_t)b * S_kv * ldkv + h * D;
  pdl_sync();
  // ---- stage K rows and V^T of this (image, head).  Only what the fragments read beyond the data is zeroed (padding
  // must be finite zeros); those regions are disjoint from the data, so one barrier serves both.
  for (int i = tid; i < (NT * 8 - S_kv) * (DP / 8); i += blockDim.x)
    *reinterpret_cast<uint4*>(Ks + (S_kv + i / (DP / 8)) * KP + (i % (DP / 8)) * 8) = make_uint4(0, 0, 0, 0);
  if constexpr (DP > D) {  // channels D .. DP - 1 = the upper half of the last 16-block: positions 4 tt + 2, + 3
    for (int i = tid; i < S_kv * 4; i += blockDim.x)
      *reinterpret_cast<uint32_t*>(Ks + (i >> 2) * KP + (DP - 16) + 4 * (i & 3) + 2) = 0u;
  }
  for (int i = tid; i < D * (NKP - S_kv); i += blockDim.x) {
    const int j = S_kv + i % (NKP - S_kv);
    Vt[(i / (NKP - S_kv)) * VP + (j & ~15) + xattn_perm16(j & 15)] = DT<T>::from_f(0.f);
  }
  const bool vec16 = (ldkv % 8 == 0) && ((reinterpret_cast<uintptr_t>(kb) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(vb) & 15) == 0);
  // consecutive threads take consecutive keys of one 8-channel chunk (V^T stores of a warp spread over the banks)
  for (int i = tid; i < S_kv * NCH; i += blockDim.x) {
    const int j = i % S_kv, c = (i / S_kv) * 8;
    uint4 kv, vv;
    const T* kp = kb + (size_t)j * ldkv + c;
    const T* vp = vb + (size_t)j * ldkv + c;
    if (vec16) {
      kv = __ldg(reinterpret_cast<const uint4*>(kp));
      vv = __ldg(reinterpret_cast<const uint4*>(vp));
    } else {
      uint32_t* k4 = reinterpret_cast<uint32_t*>(&kv);
      uint32_t* v4 = reinterpret_cast<uint32_t*>(&vv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        k4[e] = *reinterpret_cast<const uint32_t*>(kp + 2 * e);
        v4[e] = *reinterpret_cast<const uint32_t*>(vp + 2 * e);
      }
    }
    // channels c .. c+7 = half hh of their 16-block: pair tt -> positions 4 tt + 2 hh, + 1
    T* krow = Ks + j * KP + (c & ~15) + 2 * ((c >> 3) & 1);
    const uint32_t* k4 = reinterpret_cast<const uint32_t*>(&kv);
#pragma unroll
    for (int tt = 0; tt < 4; ++tt) *reinterpret_cast<uint32_t*>(krow + 4 * tt) = k4[tt];
    const T* ve = reinterpret_cast<const T*>(&vv);
    const int jp = (j & ~15) + xattn_perm16(j & 15);
#pragma unroll
    for (int e = 0; e < 8; ++e) Vt[(c + e) * VP + jp] = ve[e];
  }
  // ---- per-warp Q / O staging tile [16][QP]; its padding columns D .. DP - 1 stay zero
  T* Qs = Vt + DP * VP + warp * 16 * QP;
  if constexpr (DP > D) {
    if (lane < 16) *reinterpret_cast<uint4*>(Qs + lane * QP + D) = make_uint4(0, 0, 0, 0);
  }
  constexpr int KS = DP / 16;
  constexpr int QV = (16 * NCH + 31) / 32;  // 16-byte vectors per lane of a 16 x D tile
  uint4 qv[QV];
  auto load_q = [&](int tile) {  // rows beyond S_q: clamped (computed, never stored)
#pragma unroll
    for (int i = 0; i < QV; ++i) {
      const int idx = lane + i * 32, r = idx / NCH, ch = idx % NCH;
      if (idx < 16 * NCH)
        qv[i] = __ldg(reinterpret_cast<const uint4*>(q + ((size_t)b * S_q + min(tile * 16 + r, S_q - 1)) * ldq + h * D + ch * 8));
    }
  };
  const int row_tiles = (S_q + 15) / 16;
  const int tile0 = blockIdx.y * tiles_per_cta;
  const int tile1 = min(row_tiles, tile0 + tiles_per_cta);
  if (tile0 + warp < tile1) load_q(tile0 + warp);  // in flight across the staging barrier
  __syncthreads();
  // no CTA barrier below this point
  const int nt_full = S_kv >> 3;  // key tiles without dead keys
  for (int tile = tile0 + warp; tile < tile1; tile += XATTN_THREADS / 32) {
    // ---- Q tile -> staging tile -> A fragments: a0 (g, c) a1 (g+8, c) a2 (g, c+8) a3 (g+8, c+8), c = 16 ks + 2t
#pragma unroll
    for (int i = 0; i < QV; ++i) {
      const int idx = lane + i * 32, r = idx / NCH, ch = idx % NCH;
      if (idx < 16 * NCH) *reinterpret_cast<uint4*>(Qs + r * QP + ch * 8) = qv[i];
    }
    __syncwarp();
    uint32_t qf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int c = ks * 16 + 2 * t;
      qf[ks][0] = *reinterpret_cast<const uint32_t*>(Qs + g * QP + c);
      qf[ks][1] = *reinterpret_cast<const uint32_t*>(Qs + (g + 8) * QP + c);
      qf[ks][2] = *reinterpret_cast<const uint32_t*>(Qs + g * QP + c + 8);
      qf[ks][3] = *reinterpret_cast<const uint32_t*>(Qs + (g + 8) * QP + c + 8);
    }
    __syncwarp();  // the staging tile is free again (it takes this tile's output below)
    // the next tile's Q vectors travel while this tile is computed
    if (tile + XATTN_THREADS / 32 < tile1) load_q(tile + XATTN_THREADS / 32);
    // ---- S = Q K^T: accumulator tile nt holds (row g | g+8, keys nt*8 + 2t, +1); B fragment (b0, b1) = one 8-byte load
    float sc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint2 bf = *reinterpret_cast<const uint2*>(Ks + (nt * 8 + g) * KP + ks * 16 + 4 * t);
        MmaOp<T>::mma(sc[nt], qf[ks], bf.x, bf.y);
      }
    }
    // ---- softmax over the whole key range, on the raw scores: p = 2^(s scale - max scale), normalised after P V
    // (quad reduction: lanes 4g .. 4g+3 share rows g, g+8).  Dead keys (>= S_kv) sit in the last tile only (EXACT) or from
    // tile S_kv / 8 on (generic): their scores become -inf, their probabilities exactly 0.
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (EXACT ? (nt == NT - 1) : (nt >= nt_full)) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (nt * 8 + 2 * t + e >= S_kv) sc[nt][e] = sc[nt][2 + e] = -INFINITY;
      }
    }
    float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mxa = fmaxf(mxa, fmaxf(sc[nt][0], sc[nt][1]));
      mxb = fmaxf(mxb, fmaxf(sc[nt][2], sc[nt][3]));
    }
    mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 1));
    mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 2));
    mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 1));
    mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 2));
    const float nma = -mxa * scale_log2, nmb = -mxb * scale_log2;
    float la = 0.f, lb = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sc[nt][e] = exp2f(fmaf(sc[nt][e], scale_log2, nma));  // exp2(-inf) = 0 for the dead keys
        sc[nt][2 + e] = exp2f(fmaf(sc[nt][2 + e], scale_log2, nmb));
        la += sc[nt][e];
        lb += sc[nt][2 + e];
      }
    }
    la += __shfl_xor_sync(0xffffffffu, la, 1);
    la += __shfl_xor_sync(0xffffffffu, la, 2);
    lb += __shfl_xor_sync(0xffffffffu, lb, 1);
    lb += __shfl_xor_sync(0xffffffffu, lb, 2);
    const float ia = 1.0f / la, ib = 1.0f / lb;
    // 16-bit-rounded probabilities (values in [0, 1]: the same relative rounding as the reference's normalised 16-bit
    // softmax output) = A fragments of P V: key tiles (2 kj, 2 kj + 1) form k-step kj; a missing odd tile is zero
    uint32_t pa[NKJ][4];
#pragma unroll
    for (int kj = 0; kj < NKJ; ++kj) {
      T2 p0 = DT<T>::from_f2(sc[2 * kj][0], sc[2 * kj][1]);
      T2 p1 = DT<T>::from_f2(sc[2 * kj][2], sc[2 * kj][3]);
      pa[kj][0] = *reinterpret_cast<uint32_t*>(&p0);
      pa[kj][1] = *reinterpret_cast<uint32_t*>(&p1);
      if (2 * kj + 1 < NT) {  // (compile-time after unrolling; the index is wrapped only to stay in bounds)
        T2 p2 = DT<T>::from_f2(sc[(2 * kj + 1) % NT][0], sc[(2 * kj + 1) % NT][1]);
        T2 p3 = DT<T>::from_f2(sc[(2 * kj + 1) % NT][2], sc[(2 * kj + 1) % NT][3]);
        pa[kj][2] = *reinterpret_cast<uint32_t*>(&p2);
        pa[kj][3] = *reinterpret_cast<uint32_t*>(&p3);
      } else {
        pa[kj][2] = pa[kj][3] = 0u;
      }
    }
    // ---- O = (P V) / l in chunks of NCHUNK column tiles, written into the staging tile
#pragma unroll
    for (int n0 = 0; n0 < NCH; n0 += NCHUNK) {
      float oc[NCHUNK][4];
#pragma unroll
      for (int n = 0; n < NCHUNK; ++n) oc[n][0] = oc[n][1] = oc[n][2] = oc[n][3] = 0.f;
#pragma unroll
      for (int kj = 0; kj < NKJ; ++kj) {
#pragma unroll
        for (int n = 0; n < NCHUNK; ++n) {
          if (n0 + n < NCH) {
            // b0 (keys kj*16 + 2t, +1; dim (n0 + n)*8 + g), b1 (keys + 8): adjacent in the permuted 16-block
            const uint2 bf = *reinterpret_cast<const uint2*>(Vt + ((n0 + n) * 8 + g) * VP + kj * 16 + 4 * t);
            MmaOp<T>::mma(oc[n], pa[kj], bf.x, bf.y);
          }
        }
      }
#pragma unroll
      for (int n = 0; n < NCHUNK; ++n) {
        if (n0 + n < NCH) {
          const int c = (n0 + n) * 8 + 2 * t;
          *reinterpret_cast<T2*>(Qs + g * QP + c) = DT<T>::from_f2(oc[n][0] * ia, oc[n][1] * ia);
          *reinterpret_cast<T2*>(Qs + (g + 8) * QP + c) = DT<T>::from_f2(oc[n][2] * ib, oc[n][3] * ib);
        }
      }
    }
    __syncwarp();
    // ---- staging tile -> out, 16 bytes per lane
#pragma unroll
    for (int i = 0; i < QV; ++i) {
      const int idx = lane + i * 32, r = idx / NCH, ch = idx % NCH;
      if (idx < 16 * NCH && tile * 16 + r < S_q)
        *reinterpret_cast<uint4*>(out + ((size_t)b * S_q + tile * 16 + r) * ldo + h * D + ch * 8) =
            *reinterpret_cast<const uint4*>(Qs + r * QP + ch * 8);
    }
    __syncwarp();  // before the next tile's Q vectors overwrite the staging tile
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Temporal attention for wide heads (the prior's motion modules: 8 heads x 256 channels; motion_module.py:294-354):
// one warp slice of LPH = d / 8 lanes per (location, head); each lane keeps one 16-byte vector of q, k, v for all F
// frames (coalesced 16-byte global loads, no shared memory), the F x F partial scores are reduced across the slice
// with shuffles, and every lane writes its 8 output channels of all F frames.  qkv [(b f hw), 3C] -> out [(b f hw), C].
// ---------------------------------------------------------------------------------------------------------------
// DV = active lanes of a slice = d / 8 (DV == LPH for power-of-two head dims; the UNet's d = 40 / 80 / 160 use
// DV = 5 / 10 / 20 of LPH = 8 / 16 / 32 lanes, the idle lanes contribute zeros to the reductions).
#ifndef RCDM_TEMPORAL_MIN_CTAS
#define RCDM_TEMPORAL_MIN_CTAS 2
#endif
template <typename T, int F, int LPH, int DV = LPH>
__global__ void __launch_bounds__(256, RCDM_TEMPORAL_MIN_CTAS)
temporal_attn_wide_kernel(const T* __restrict__ qkv, T* __restrict__ out, int batch, int hw, int heads, float scale) {
  constexpr int d = DV * 8;
  const int C = heads * d, ld = 3 * C;
  const size_t slice = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / LPH;  // (b, pix, head)
  const int sub = threadIdx.x % LPH;
  pdl_sync();
  const size_t total = (size_t)batch * hw * heads;
  const bool active = slice < total;
  const size_t sl = active ? slice : total - 1;  // inactive lanes shadow a valid slice so that shuffles stay full-warp
  const int h = (int)(sl % heads);
  const size_t loc = sl / heads;
  const int pix = (int)(loc % hw);
  const int b = (int)(loc / hw);
  const size_t row0 = (size_t)b * F * hw + pix;
  const bool lane_on = DV == LPH || sub < DV;
  float q[F][8], k[F][8], v[F][8];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const T* p = qkv + (row0 + (size_t)f * hw) * ld + h * d + sub * 8;
    const uint4 z = make_uint4(0, 0, 0, 0);
    unpack8<T>(lane_on ? __ldg(reinterpret_cast<const uint4*>(p)) : z, q[f]);
    unpack8<T>(lane_on ? __ldg(reinterpret_cast<const uint4*>(p + C)) : z, k[f]);
    unpack8<T>(lane_on ? __ldg(reinterpret_cast<const uint4*>(p + 2 * C)) : z, v[f]);
  }
  float s[F][F];
#pragma unroll
  for (int i = 0; i < F; ++i)
#pragma unroll
    for (int j = 0; j < F; ++j) {
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) a = fmaf(q[i][e], k[j][e], a);
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      s[i][j] = a * scale;
    }
#pragma unroll
  for (int i = 0; i < F; ++i) {
    float mx = s[i][0];
#pragma unroll
    for (int j = 1; j < F; ++j) mx = fmaxf(mx, s[i][j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      s[i][j] = __expf(s[i][j] - mx);
      sum += s[i][j];
    }
    const float inv = 1.0f / sum;
    float o8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o8[e] = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      const float pj = s[i][j] * inv;
#pragma unroll
      for (int e = 0; e < 8; ++e) o8[e] = fmaf(pj, v[j][e], o8[e]);
    }
    if (active && lane_on)
      *reinterpret_cast<uint4*>(out + (row0 + (size_t)i * hw) * C + h * d + sub * 8) = pack8<T>(o8);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Per-step token assembly (myprior_transformer.py:335-384): the token matrix x[(b, s), C] is the step-invariant part
// `base` (text tokens, the three projected conditioning embeddings, the prd token; positional embedding already
// added) with two rows per sample rewritten every step:
//   row t_row = round(time_embedding[step] + pos[t_row]),  row h_row = round(proj_in(latents)[b % n_lat] + pos[h_row])
// (`torch.cat([latents] * 2)` for CFG, prior_pipeline.py:303, is the modulo).  `step` is a device counter so that the
// launch can be replayed from a CUDA graph.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void prior_assemble_kernel(const T* __restrict__ base, const T* __restrict__ temb_tab,
                                      const T* __restrict__ hproj, const T* __restrict__ pos, T* __restrict__ x,
                                      int B, int S, int C, int t_row, int h_row, int n_lat,
                                      const int* __restrict__ step) {
  pdl_sync();
  const int st = step ? *step : 0;
  const int cv = C / 8;
  const size_t total = (size_t)B * S * cv;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % cv);
    const int s = (int)((idx / cv) % S);
    const int b = (int)(idx / ((size_t)cv * S));
    uint4 val;
    if (s == t_row || s == h_row) {
      const T* src = s == t_row ? temb_tab + (size_t)st * C : hproj + (size_t)(b % n_lat) * C;
      float a[8], p8[8];
      unpack8<T>(*reinterpret_cast<const uint4*>(src + v * 8), a);
      unpack8<T>(*reinterpret_cast<const uint4*>(pos + (size_t)s * C + v * 8), p8);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += p8[i];
      val = pack8<T>(a);
    } else {
      val = *reinterpret_cast<const uint4*>(base + ((size_t)b * S + s) * C + v * 8);
    }
    *reinterpret_cast<uint4*>(x + ((size_t)b * S + s) * C + v * 8) = val;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused classifier-free guidance + UnCLIPScheduler.step (prior_pipeline.py:316-333; diffusers 0.24.0
// scheduling_unclip.step, restated in oracle/diffusers_restated.py).  Every torch op of the reference rounds its
// result to the sample dtype; the kernel rounds at the same points, so given the same prediction and noise the
// result is bit-identical to the python loop:
//   pred = pu + g * (pt - pu);  x0 = pred | (x - eps_scale * pred) / eps_div;  x0 = clamp(x0, -clip, clip)
//   x' = c_x0 * x0 + c_x * x (+ sigma * noise[step] when sigma > 0)
// coef: float[steps][8] = {c_x0, c_x, sigma, eps_scale, eps_div, clip (<= 0: no clipping), epsilon?, -}.
// Single CTA (n = frames * D is a few thousand elements); the last thread to finish advances the step counter.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
unclip_cfg_step_kernel(const T* __restrict__ pred, T* __restrict__ latents, const T* __restrict__ noise_tab,
                       const float* __restrict__ coef, int n, int do_cfg, float guidance, int* __restrict__ step,
                       int advance) {
  pdl_sync();
  const int st = step ? *step : 0;
  const float* cf = coef + (size_t)st * 8;
  const float c_x0 = cf[0], c_x = cf[1], sigma = cf[2], eps_scale = cf[3], eps_div = cf[4], clip = cf[5];
  const bool eps_mode = cf[6] != 0.f;
  auto rn = [](float v) { return DT<T>::to_f(DT<T>::from_f(v)); };
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float p;
    if (do_cfg) {
      const float pu = DT<T>::to_f(pred[i]), pt = DT<T>::to_f(pred[n + i]);
      p = rn(pu + rn(guidance * rn(pt - pu)));
    } else {
      p = DT<T>::to_f(pred[i]);
    }
    const float x = DT<T>::to_f(latents[i]);
    float x0 = p;
    if (eps_mode) x0 = rn(rn(x - rn(eps_scale * p)) * __frcp_rn(eps_div));  // torch divides by a scalar as x * (1 / s)
    if (clip > 0.f) x0 = fminf(fmaxf(x0, -clip), clip);
    float prev = rn(rn(c_x0 * x0) + rn(c_x * x));
    if (sigma > 0.f) {
      const float nz = DT<T>::to_f(noise_tab[(size_t)st * n + i]);
      prev = rn(prev + rn(sigma * nz));
    }
    latents[i] = DT<T>::from_f(prev);
  }
  __syncthreads();
  if (advance && step && threadIdx.x == 0) *step = st + 1;
}

}  // namespace rcdm
