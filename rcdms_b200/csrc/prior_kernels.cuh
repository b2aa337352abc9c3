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
