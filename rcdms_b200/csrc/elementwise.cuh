// Small HBM-bound kernels around the tensor-core path: boundary layout conversion (NCFHW <-> channels-last
// tokens), conv_in im2col (9 input channels), nearest 2x upsample, the timestep-embedding MLP, the fused
// CFG + DDIM step, and the weight (re)packing kernels.
#pragma once
#include "common.cuh"

namespace rcdm {

__device__ __forceinline__ float load_any(const void* p, int dt, size_t i) {
  if (dt == DT_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == DT_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void store_any(void* p, int dt, size_t i, float v) {
  if (dt == DT_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == DT_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// conv_in (unet.py:403; InflatedConv3d resnet.py:10-18): sample (b, Cin, f, h, w) in any dtype ->
// im2col matrix A[(b f y x), Kpad] with k = (ky*3 + kx)*Cin + c (zero for padding taps and k >= 9*Cin).
// The sample is first rounded to the compute dtype, as `.to(dtype=latents_dtype)` does (RCDMs_pipeline.py:486).
template <typename T>
__global__ void im2col_in_kernel(const void* __restrict__ x, int x_dt, T* __restrict__ A, int B, int Cin, int F, int H,
                                 int W, int Kpad) {
  const size_t total = (size_t)B * F * H * W * Kpad;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kpad);
    size_t m = idx / Kpad;
    const int xx = (int)(m % W);
    m /= W;
    const int yy = (int)(m % H);
    m /= H;
    const int f = (int)(m % F);
    const int b = (int)(m / F);
    float v = 0.f;
    if (k < 9 * Cin) {
      const int tap = k / Cin, c = k % Cin;
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      if (sy >= 0 && sy < H && sx >= 0 && sx < W)
        v = load_any(x, x_dt, ((((size_t)b * Cin + c) * F + f) * H + sy) * W + sx);
    }
    A[idx] = DT<T>::from_f(v);
  }
}

// tokens [(b f y x), C] -> (b, C, f, h, w) in the caller's dtype (conv_out result, unet.py:457-460)
template <typename T>
__global__ void tokens_to_ncfhw_kernel(const T* __restrict__ tok, void* __restrict__ out, int out_dt, int B, int C,
                                       int F, int HW) {
  const size_t total = (size_t)B * C * F * HW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(idx % HW);
    size_t r = idx / HW;
    const int f = (int)(r % F);
    r /= F;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    store_any(out, out_dt, idx, DT<T>::to_f(tok[(((size_t)b * F + f) * HW + pix) * C + c]));
  }
}

// nearest-neighbour 2x spatial upsample on channels-last images (Upsample3D, resnet.py:65)
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
  const int vecs = C / 8;
  const size_t total = (size_t)N * 2 * H * 2 * W * vecs;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vecs);
    size_t r = idx / vecs;
    const int ox = (int)(r % (2 * W));
    r /= 2 * W;
    const int oy = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + oy / 2) * W + ox / 2) * C + v * 8));
    *reinterpret_cast<uint4*>(y + idx * 8) = val;
  }
}

// 3x3 / pad 1 / stride 1 convolution for a SMALL input channel count (VAE conv_in: 4 -> 512 at 64x64, 3 -> 128 at 512x512;
// < 0.2 % of the VAE's FLOPs): CUDA cores, one thread per (pixel, 8 output channels), fp32 accumulation.
// x [n,h,w,cin] channels-last, w packed [cout, 9*cin] tap-major (rcdm_pack_conv3x3), out [n,h,w,cout].
template <typename T>
__global__ void conv3x3_small_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                     T* __restrict__ out, int N, int H, int W, int Cin, int Cout) {
  const int og = Cout / 8;
  const size_t total = (size_t)N * H * W * og;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % og);
    size_t pix = idx / og;
    const int xx = (int)(pix % W);
    const int yy = (int)((pix / W) % H);
    const int n = (int)(pix / ((size_t)W * H));
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = bias ? __ldg(bias + g * 8 + o) : 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      const T* src = x + (((size_t)n * H + sy) * W + sx) * Cin;
      for (int c = 0; c < Cin; ++c) {
        const float xv = DT<T>::to_f(src[c]);
#pragma unroll
        for (int o = 0; o < 8; ++o)
          acc[o] = fmaf(xv, DT<T>::to_f(__ldg(w + (size_t)(g * 8 + o) * 9 * Cin + tap * Cin + c)), acc[o]);
      }
    }
    *reinterpret_cast<uint4*>(out + pix * Cout + g * 8) = pack8<T>(acc);
  }
}

// out[m, n] = sum_k x[m, k] w[n, k] + bias[n] for tiny N, K (<= 16: the VAE's 1x1 quant_conv / post_quant_conv on 8 / 4
// latent channels); one thread per row.
template <typename T>
__global__ void linear_small_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                    T* __restrict__ out, size_t M, int N, int K) {
  for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (size_t)gridDim.x * blockDim.x) {
    float xv[16];
    for (int k = 0; k < K; ++k) xv[k] = DT<T>::to_f(x[m * K + k]);
    for (int n = 0; n < N; ++n) {
      float acc = bias ? __ldg(bias + n) : 0.f;
      for (int k = 0; k < K; ++k) acc = fmaf(xv[k], DT<T>::to_f(__ldg(w + n * K + k)), acc);
      out[m * N + n] = DT<T>::from_f(acc);
    }
  }
}

// row softmax in place: x[r, 0..cols) <- softmax(scale * x[r, :]) (fp32 inside, one rounding out); one CTA per row.
// The VAE's single-head d = 512 attention over 4096 tokens (diffusers Attention in AutoencoderKL's mid block) is
// two tensor-core GEMMs around this kernel.
template <typename T>
__global__ void softmax_rows_kernel(T* __restrict__ x, int cols, int ld, float scale) {
  __shared__ float red[32];
  T* row = x + (size_t)blockIdx.x * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int c = threadIdx.x * 8; c < cols; c += blockDim.x * 8) {
    float f[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) mx = fmaxf(mx, f[i]);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < nw; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  const float l2e = 1.4426950408889634f * scale;
  for (int c = threadIdx.x * 8; c < cols; c += blockDim.x * 8) {
    float f[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += exp2f((f[i] - mx) * l2e);
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < nw; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x * 8; c < cols; c += blockDim.x * 8) {
    float f[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = exp2f((f[i] - mx) * l2e) * inv;
    *reinterpret_cast<uint4*>(row + c) = pack8<T>(f);
  }
}

// ---- timestep embedding (unet.py:367-389 + resnet.py:190-191), all in fp32 from 16-bit weights ----------
// step 1: sinusoid (flip_sin_to_cos, freq_shift 0) -> rounded to T (".to(dtype)") -> linear_1 -> SiLU
// one warp per output feature.  t comes from device memory (int64) when t_dev != nullptr.
template <typename T>
__global__ void temb_linear1_kernel(const int64_t* t_dev, float t_host, const T* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ out, int c0, int n_out,
                                    int flip_sin_to_cos, float freq_shift) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (o >= n_out) return;
  const float t = t_dev ? (float)(*t_dev) : t_host;
  const int half = c0 / 2;
  float acc = 0.f;
  for (int i = lane; i < c0; i += 32) {
    // column i of the embedding: [cos | sin] when flipped, [sin | cos] otherwise
    const int fi = i % half;
    const bool is_sin = flip_sin_to_cos ? (i >= half) : (i < half);
    const float freq = expf(-logf(10000.0f) * (float)fi / ((float)half - freq_shift));
    const float arg = t * freq;
    float e = is_sin ? sinf(arg) : cosf(arg);
    e = DT<T>::to_f(DT<T>::from_f(e));
    acc += e * DT<T>::to_f(w[(size_t)o * c0 + i]);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) out[o] = silu_f(acc + bias[o]);
}

// generic fp32-activation GEMV: out[o] = act_out(dot(x, w[o,:]) + bias[o]); one warp per output
template <typename T>
__global__ void gemv_kernel(const float* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                            const float* __restrict__ bias2, float* __restrict__ out, int n_in, int n_out,
                            int silu_out) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (o >= n_out) return;
  float acc = 0.f;
  for (int i = lane * 8; i < n_in; i += 256) {
    float f[8];
    unpack8<T>(__ldg(reinterpret_cast<const uint4*>(w + (size_t)o * n_in + i)), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += f[e] * x[i + e];
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    acc += bias[o];
    if (bias2) acc += bias2[o];
    out[o] = silu_out ? silu_f(acc) : acc;
  }
}

// ---- fused classifier-free guidance + DDIM step + next UNet input (RCDMs_pipeline.py:482-497) ------------
// eps:      (2B, 4, f, h, w)   [uncond clips | cond clips]   (B, ...) when !cfg
// latents:  (B, 4, f, h, w)  updated in place (fp32 master copy + optional user-dtype copy)
// next_in:  (2B, 9, f, h, w)   [latents | mask | masked_latents] for both CFG halves (compute dtype T)
// coef: {1/sqrt(abar_t), sqrt(1-abar_t), sqrt(abar_prev), sqrt(1-abar_prev)} read from a device table at *step_idx
// when table != nullptr (graph replay), else passed by value.
struct DdimArgs {
  const void* eps;
  int eps_dt;
  float* latents;        // fp32 master [B,4,f,h,w]
  void* latents_out;     // optional copy in user dtype
  int latents_out_dt;
  void* next_in;         // (2B or B, 9, f, h, w) in next_dt, may be nullptr
  int next_dt;
  const void* mask;      // (B,1,f,h,w) any dtype
  int mask_dt;
  const void* masked;    // (B,4,f,h,w)
  int masked_dt;
  int B, FHW;            // clips, f*h*w
  int cfg;
  float guidance;
  float c[4];
  const float4* table;   // [steps] or nullptr
  const int* step_idx;   // device counter (read)
  int round_dt;          // dtype the reference would hold latents in (rounding point), DT_*
};

__device__ __forceinline__ float round_to(float v, int dt) {
  if (dt == DT_F16) return __half2float(__float2half_rn(v));
  if (dt == DT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  return v;
}

// Arithmetic mirrors the reference op by op: torch evaluates every elementwise op on 16-bit tensors in fp32 and
// rounds the result to the tensor dtype, scalars (python floats / 0-dim fp32 CPU tensors) stay fp32, and a division
// by a CPU scalar is a multiplication by its fp32 reciprocal.  So, with r() = round to the latents dtype:
//   CFG  (RCDMs_pipeline.py:493-494):  e = r(eu + r(g * r(ec - eu)))
//   DDIM (diffusers scheduling_ddim.step, eta = 0, epsilon prediction, no clipping):
//        x0 = r(r(x - r(c1 * e)) * c0inv) ;  x_prev = r(r(c2 * x0) + r(c3 * e))
// __fmul_rn / __fadd_rn keep nvcc from contracting across the reference's rounding points (fp32 latents too).
static __global__ void ddim_cfg_step_kernel(const DdimArgs a) {
  const size_t n = (size_t)a.B * 4 * a.FHW;
  float4 co = make_float4(a.c[0], a.c[1], a.c[2], a.c[3]);
  if (a.table) co = a.table[*a.step_idx];
  const int rd = a.round_dt;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(idx % a.FHW);
    const int ch = (int)((idx / a.FHW) % 4);
    const int b = (int)(idx / ((size_t)4 * a.FHW));
    float e;
    if (a.cfg) {
      const float eu = load_any(a.eps, a.eps_dt, idx);
      const float ec = load_any(a.eps, a.eps_dt, idx + n);
      const float d = round_to(__fadd_rn(ec, -eu), rd);
      e = round_to(__fadd_rn(eu, round_to(__fmul_rn(a.guidance, d), rd)), rd);
    } else {
      e = load_any(a.eps, a.eps_dt, idx);
    }
    const float x = a.latents[idx];
    const float t2 = round_to(__fadd_rn(x, -round_to(__fmul_rn(co.y, e), rd)), rd);
    const float x0 = round_to(__fmul_rn(t2, co.x), rd);
    const float xn = round_to(__fadd_rn(round_to(__fmul_rn(co.z, x0), rd), round_to(__fmul_rn(co.w, e), rd)), rd);
    a.latents[idx] = xn;
    if (a.latents_out) store_any(a.latents_out, a.latents_out_dt, idx, xn);
    if (a.next_in) {
      const int reps = a.cfg ? 2 : 1;
      for (int r = 0; r < reps; ++r) {
        const size_t ob = ((size_t)(r * a.B + b) * 9) * a.FHW;
        store_any(a.next_in, a.next_dt, ob + (size_t)ch * a.FHW + pix, xn);
        store_any(a.next_in, a.next_dt, ob + (size_t)(5 + ch) * a.FHW + pix,
                  load_any(a.masked, a.masked_dt, idx));
        if (ch == 0)
          store_any(a.next_in, a.next_dt, ob + (size_t)4 * a.FHW + pix,
                    load_any(a.mask, a.mask_dt, (size_t)b * a.FHW + pix));
      }
    }
  }
}

static __global__ void advance_step_kernel(int* step_idx) { *step_idx += 1; }

// ---- weight packing ------------------------------------------------------------------------------------
// dst[(row_map(n)) * ldd + col_off + k'] for src viewed as [N, K] (K = Cin*taps for convs)
// kind 0: matrix [N,K] row-major;  kind 1: conv3x3 [N, Cin, 3, 3] -> k' = tap*Cin + c
// geglu_bn > 0: GEGLU row interleave for tile width geglu_bn (rows [0,N/2) = h, [N/2,N) = gate)
__device__ __forceinline__ int geglu_row(int r, int N, int bn) {
  const int half = N / 2, hb = bn / 2;
  const int gate = r >= half;
  const int j = gate ? r - half : r;
  return (j / hb) * bn + (gate ? hb : 0) + (j % hb);
}

template <typename T>
__global__ void pack_weight_kernel(const void* __restrict__ src, int src_dt, T* __restrict__ dst, int N, int K, int ldd,
                                   int col_off, int row_off, int kind, int Cin, int geglu_bn) {
  const size_t total = (size_t)N * K;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / K);
    const int k = (int)(idx % K);
    int kk = k;
    if (kind == 1) {  // src k = c*9 + tap
      const int c = k / 9, tap = k % 9;
      kk = tap * Cin + c;
    }
    const int row = (geglu_bn > 0 ? geglu_row(n, N, geglu_bn) : n) + row_off;
    dst[(size_t)row * ldd + col_off + kk] = DT<T>::from_f(load_any(src, src_dt, idx));
  }
}

static __global__ void pack_vec_kernel(const void* __restrict__ src, int src_dt, float* __restrict__ dst, int N, int off,
                                int geglu_bn, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int row = (geglu_bn > 0 ? geglu_row(i, N, geglu_bn) : i) + off;
  const float v = load_any(src, src_dt, i);
  dst[row] = accumulate ? dst[row] + v : v;
}

// Upsample3D folded into its conv (gemm_tcgen05.cuh, SEG_UP2): for output parity class (py, px) the 3x3 taps that land on
// the same input pixel are summed (in fp32, one rounding): rows S(0,0) = {0}, S(0,1) = {1,2}, S(1,0) = {0,1}, S(1,1) = {2}.
// w [Cout, 9*Cin] tap-major (packed conv layout) -> wf [4 classes][Cout][4*Cin] with k = (ty*2 + tx)*Cin + c.
template <typename T>
__global__ void fold_upsample_kernel(const T* __restrict__ w, T* __restrict__ wf, int Cout, int Cin) {
  const size_t total = (size_t)4 * Cout * 4 * Cin;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cin);
    const int t2 = (int)((idx / Cin) % 4);
    const int co = (int)((idx / ((size_t)4 * Cin)) % Cout);
    const int cls = (int)(idx / ((size_t)4 * Cin * Cout));
    const int py = cls >> 1, px = cls & 1, ty = t2 >> 1, tx = t2 & 1;
    const int ky0 = py == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), ky1 = py == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
    const int kx0 = px == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), kx1 = px == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
    float acc = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) acc += DT<T>::to_f(w[(size_t)co * 9 * Cin + (ky * 3 + kx) * Cin + c]);
    wf[idx] = DT<T>::from_f(acc);
  }
}

// Whole-state-dict packing in ONE launch: a device table of jobs (one per state-dict entry), each owning a contiguous
// range of CTAs; a CTA finds its job by binary search over the ranges' first block index.
struct PackJob {
  const void* src;
  void* dst;       // T* (matrix / conv kinds) or float* (vector kind)
  int src_dt;
  int kind;        // 0 matrix [N,K], 1 conv3x3 [N,Cin,3,3] -> tap-major, 2 fp32 vector [N]
  int N, K, ldd, col_off, row_off, Cin, geglu_bn;
  int block0;      // first CTA of this job
  int nblocks;
};
template <typename T>
__global__ void pack_many_kernel(const PackJob* __restrict__ jobs, int njobs) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {  // last job with block0 <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block0 <= (int)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const PackJob j = jobs[lo];
  const size_t total = (size_t)j.N * (j.kind == 2 ? 1 : j.K);
  const size_t stride = (size_t)j.nblocks * blockDim.x;
  for (size_t idx = (size_t)(blockIdx.x - j.block0) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const float v = load_any(j.src, j.src_dt, idx);
    if (j.kind == 2) {
      const int row = (j.geglu_bn > 0 ? geglu_row((int)idx, j.N, j.geglu_bn) : (int)idx) + j.row_off;
      reinterpret_cast<float*>(j.dst)[row] = v;
    } else {
      const int n = (int)(idx / j.K);
      const int k = (int)(idx % j.K);
      int kk = k;
      if (j.kind == 1) {  // src k = c*9 + tap
        const int c = k / 9, tap = k % 9;
        kk = tap * j.Cin + c;
      }
      const int row = (j.geglu_bn > 0 ? geglu_row(n, j.N, j.geglu_bn) : n) + j.row_off;
      reinterpret_cast<T*>(j.dst)[(size_t)row * j.ldd + j.col_off + kk] = DT<T>::from_f(v);
    }
  }
}

// LayerNorm fold (see GemmParams): one warp per weight row n.
//   wf[n,k] = W[n,k] * gamma[k] - mean_k(W[n,:] * gamma)   (rounded to T; centred rows absorb the "- mean * u" term)
//   c[f][n] = sum_k (beta[k] + pe[f][k]) * W[n,k] + bias[n]        (pe / bias optional; frames >= 1)
template <typename T>
__global__ void fold_ln_kernel(const T* __restrict__ W, T* __restrict__ Wf, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ pe,
                               const float* __restrict__ bias, float* __restrict__ c, int N, int K, int frames) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float sg = 0.f, sb = 0.f, sp[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < K; k += 32) {
    const float w = DT<T>::to_f(W[(size_t)n * K + k]);
    sg = fmaf(w, gamma[k], sg);
    sb = fmaf(beta[k], w, sb);
    if (pe) {
#pragma unroll
      for (int f = 0; f < 5; ++f)
        if (f < frames) sp[f] = fmaf(pe[(size_t)f * K + k], w, sp[f]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
#pragma unroll
    for (int f = 0; f < 5; ++f) sp[f] += __shfl_xor_sync(0xffffffffu, sp[f], o);
  }
  const float centre = sg / (float)K;
  for (int k = lane; k < K; k += 32)
    Wf[(size_t)n * K + k] = DT<T>::from_f(fmaf(DT<T>::to_f(W[(size_t)n * K + k]), gamma[k], -centre));
  if (lane == 0) {
    const float b = bias ? bias[n] : 0.f;
    for (int f = 0; f < frames; ++f) c[(size_t)f * N + n] = sb + sp[f < 5 ? f : 0] + b;
  }
}

// per-row (sum, sum of squares) of a [rows, C] matrix in the single-part layout stats[0][rows] (stand-alone entry
// point rcdm_linear_ln; inside the UNet the producing GEMM's epilogue emits the statistics)
template <typename T>
__global__ void rowstats_kernel(const T* __restrict__ x, float2* __restrict__ stats, int rows, int C) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float s = 0.f, ss = 0.f;
  for (int k = lane; k < C; k += 32) {
    const float v = DT<T>::to_f(x[(size_t)r * C + k]);
    s += v;
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) stats[r] = make_float2(s, ss);
}

template <typename T>
__global__ void cast_rows_kernel(const void* __restrict__ src, int src_dt, T* __restrict__ dst, size_t n) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x)
    dst[idx] = DT<T>::from_f(load_any(src, src_dt, idx));
}

}  // namespace rcdm
