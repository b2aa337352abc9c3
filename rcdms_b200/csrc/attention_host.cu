// Host side of the attention kernels: 5-D TMA maps over the fused projection outputs, launch, and a
// CUDA-core debug kernel with identical semantics (RCDM_SIMPLE=1 only).
#include "launch.h"
#include "prior_kernels.cuh"

namespace rcdm {
int num_sms();  // gemm_host.cu
}

namespace rcdm {

static int pick_dpad(int d) {
  if (d <= 16) return 16;
  if (d <= 32) return 32;
  if (d <= 48) return 48;
  if (d <= 80) return 80;
  if (d <= 160) return 160;
  return -1;
}

// (8 elems, rows-per-image, 16-byte chunks of the head, heads, images); box (8, 128 | 64, dpad/8, 1, 1)
static bool head_map(CUtensorMap* m, const void* base, int ld, int S, int heads, int d, int batch, int dpad,
                     int box_rows, std::string* err) {
  uint64_t dims[5] = {8, (uint64_t)S, (uint64_t)d / 8, (uint64_t)heads, (uint64_t)batch};
  uint64_t str[4] = {(uint64_t)ld * 2, 16, (uint64_t)d * 2, (uint64_t)S * ld * 2};
  uint32_t box[5] = {8, (uint32_t)box_rows, (uint32_t)dpad / 8, 1, 1};
  return encode_tmap(m, base, 5, dims, str, box, false, err);
}

static bool short_kv_dim(int d) { return d == 8 || d == 16 || d == 32 || d == 40 || d == 64 || d == 80 || d == 160; }

bool attn_prepare(const AttnDesc& d, AttnLaunch* l, std::string* err) {
  const int dpad = pick_dpad(d.d);
  if (dpad < 0 || d.d % 8 != 0) {
    if (err) *err = "attn_prepare: unsupported head dim";
    return false;
  }
  memset(&l->maps, 0, sizeof l->maps);
  l->dpad = dpad;
  l->dt = d.dt;
  l->short_kv = 0;
  l->desc = d;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (opt(OPT_ATTN_SHORT_KV) && d.S_kv <= XATTN_SK && short_kv_dim(d.d) && d.ldq % 8 == 0 && d.ldkv % 2 == 0 &&
      d.ldo % 8 == 0 && al16(d.q) && al16(d.out)) {
    // short key range: one warp per 16 query rows, whole key range in registers (cross_attn_mma_kernel)
    const int row_tiles = (d.S_q + 15) / 16, pairs = d.batch * d.heads;
    int chunks = (4 * num_sms() + pairs - 1) / pairs;            // ~4 CTAs per SM in flight
    const int max_chunks = (row_tiles + 3) / 4;                  // at least one tile per warp
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    l->tiles_per_cta = (row_tiles + chunks - 1) / chunks;
    l->grid = dim3(pairs, (row_tiles + l->tiles_per_cta - 1) / l->tiles_per_cta, 1);
    l->p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)d.d));
    l->short_kv = 1;
    return true;
  }
  if (!head_map(&l->maps.q, d.q, d.ldq, d.S_q, d.heads, d.d, d.batch, dpad, 128, err)) return false;
  if (!head_map(&l->maps.k, d.k, d.ldkv, d.S_kv, d.heads, d.d, d.batch, dpad, 64, err)) return false;
  // V: only the real 16-byte chunks when the head dim leaves a padded chunk (the kernel presets that chunk, see MSUM)
  if (!head_map(&l->maps.v, d.v, d.ldkv, d.S_kv, d.heads, d.d, d.batch, d.d < dpad ? d.d : dpad, 64, err)) return false;
  l->p.S_q = d.S_q;
  l->p.S_kv = d.S_kv;
  l->p.heads = d.heads;
  l->p.d = d.d;
  l->p.batch = d.batch;
  l->p.out = d.out;
  l->p.ldo = d.ldo;
  l->p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)d.d));
  l->grid = dim3((d.S_q + 127) / 128, d.heads, d.batch);
  return true;
}

template <typename T, int DPAD> static void launch_one(const AttnLaunch& l, cudaStream_t s) {
  if (l.p.d < DPAD)
    launch_k(flash_attn4_kernel<T, DPAD, true>, l.grid, dim3(Attn4Cfg<DPAD>::THREADS), Attn4Cfg<DPAD>::SMEM_BYTES, s, l.maps, l.p);
  else
    launch_k(flash_attn4_kernel<T, DPAD, false>, l.grid, dim3(Attn4Cfg<DPAD>::THREADS), Attn4Cfg<DPAD>::SMEM_BYTES, s, l.maps, l.p);
}
template <typename T> static void launch_dt(const AttnLaunch& l, cudaStream_t s) {
  switch (l.dpad) {
    case 16: launch_one<T, 16>(l, s); break;
    case 32: launch_one<T, 32>(l, s); break;
    case 48: launch_one<T, 48>(l, s); break;
    case 80: launch_one<T, 80>(l, s); break;
    default: launch_one<T, 160>(l, s); break;
  }
}
template <typename T, int D, int NT, bool EXACT> static void launch_short(const AttnLaunch& l, cudaStream_t s) {
  const AttnDesc& d = l.desc;
  launch_k(cross_attn_mma_kernel<T, D, NT, EXACT>, l.grid, dim3(XATTN_THREADS), xattn_smem_bytes(D, NT), s,
           reinterpret_cast<const T*>(d.q), d.ldq, reinterpret_cast<const T*>(d.k), reinterpret_cast<const T*>(d.v), d.ldkv,
           reinterpret_cast<T*>(d.out), d.ldo, d.S_q, d.S_kv, d.heads, l.tiles_per_cta, l.p.scale_log2);
}
// the UNet's head dims with the exact key-tile counts of its key ranges (64 keys: 8x8 self-attention; 85 / 91 context tokens)
template <typename T, int D> static void launch_short_unet(const AttnLaunch& l, cudaStream_t s) {
  switch ((l.desc.S_kv + 7) / 8) {
    case 8: launch_short<T, D, 8, true>(l, s); break;
    case 11: launch_short<T, D, 11, true>(l, s); break;
    case 12: launch_short<T, D, 12, true>(l, s); break;
    default: launch_short<T, D, XATTN_NT, false>(l, s); break;
  }
}
template <typename T> static void launch_short_dt(const AttnLaunch& l, cudaStream_t s) {
  switch (l.desc.d) {
    case 8: launch_short<T, 8, XATTN_NT, false>(l, s); break;
    case 16: launch_short<T, 16, XATTN_NT, false>(l, s); break;
    case 32: launch_short<T, 32, XATTN_NT, false>(l, s); break;
    case 40: launch_short_unet<T, 40>(l, s); break;
    case 64: launch_short<T, 64, XATTN_NT, false>(l, s); break;
    case 80: launch_short_unet<T, 80>(l, s); break;
    default: launch_short_unet<T, 160>(l, s); break;
  }
}
void attn_launch(const AttnLaunch& l, cudaStream_t s) {
  if (l.short_kv) {
    if (l.dt == DT_F16) launch_short_dt<__half>(l, s);
    else launch_short_dt<__nv_bfloat16>(l, s);
    return;
  }
  if (l.dt == DT_F16) launch_dt<__half>(l, s);
  else launch_dt<__nv_bfloat16>(l, s);
}

template <typename T, int DPAD> static cudaError_t set_attr() {
  cudaError_t e = cudaFuncSetAttribute(flash_attn4_kernel<T, DPAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Attn4Cfg<DPAD>::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(flash_attn4_kernel<T, DPAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Attn4Cfg<DPAD>::SMEM_BYTES);
  return e;
}
template <typename T> static cudaError_t set_attr_dt() {
  cudaError_t e = set_attr<T, 16>();
  if (e == cudaSuccess) e = set_attr<T, 32>();
  if (e == cudaSuccess) e = set_attr<T, 48>();
  if (e == cudaSuccess) e = set_attr<T, 80>();
  if (e == cudaSuccess) e = set_attr<T, 160>();
  // cross_attn_mma_kernel: the instantiations that stage more than 48 KB (K rows + V^T + the warps' Q / O tiles)
  auto big = [&](auto kernel, size_t bytes) {
    if (e == cudaSuccess && bytes > 48 * 1024)
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  };
  big(cross_attn_mma_kernel<T, 64, XATTN_NT, false>, xattn_smem_bytes(64, XATTN_NT));
  big(cross_attn_mma_kernel<T, 80, 8, true>, xattn_smem_bytes(80, 8));
  big(cross_attn_mma_kernel<T, 80, 11, true>, xattn_smem_bytes(80, 11));
  big(cross_attn_mma_kernel<T, 80, 12, true>, xattn_smem_bytes(80, 12));
  big(cross_attn_mma_kernel<T, 80, XATTN_NT, false>, xattn_smem_bytes(80, XATTN_NT));
  big(cross_attn_mma_kernel<T, 160, 8, true>, xattn_smem_bytes(160, 8));
  big(cross_attn_mma_kernel<T, 160, 11, true>, xattn_smem_bytes(160, 11));
  big(cross_attn_mma_kernel<T, 160, 12, true>, xattn_smem_bytes(160, 12));
  big(cross_attn_mma_kernel<T, 160, XATTN_NT, false>, xattn_smem_bytes(160, XATTN_NT));
  return e;
}
bool attn_setup_attributes(std::string* err) {
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = set_attr_dt<__half>();
  if (e == cudaSuccess) e = set_attr_dt<__nv_bfloat16>();
  if (e != cudaSuccess) {
    if (err) *err = std::string("cudaFuncSetAttribute(attn): ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

// ---- debug kernel: one thread per (image, head, query row) --------------------------------------------
template <typename T>
__global__ void attn_simple_kernel(const AttnDesc a) {
  const size_t total = (size_t)a.batch * a.heads * a.S_q;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int i = (int)(idx % a.S_q);
  const int h = (int)((idx / a.S_q) % a.heads);
  const int b = (int)(idx / ((size_t)a.S_q * a.heads));
  const T* q = reinterpret_cast<const T*>(a.q) + ((size_t)b * a.S_q + i) * a.ldq + h * a.d;
  const T* kb = reinterpret_cast<const T*>(a.k) + (size_t)b * a.S_kv * a.ldkv + h * a.d;
  const T* vb = reinterpret_cast<const T*>(a.v) + (size_t)b * a.S_kv * a.ldkv + h * a.d;
  const float scale = rsqrtf((float)a.d);
  float mx = -INFINITY;
  for (int j = 0; j < a.S_kv; ++j) {
    float s = 0.f;
    for (int c = 0; c < a.d; ++c) s += DT<T>::to_f(q[c]) * DT<T>::to_f(kb[(size_t)j * a.ldkv + c]);
    mx = fmaxf(mx, s * scale);
  }
  float o[160];
  for (int c = 0; c < a.d; ++c) o[c] = 0.f;
  float l = 0.f;
  for (int j = 0; j < a.S_kv; ++j) {
    float s = 0.f;
    for (int c = 0; c < a.d; ++c) s += DT<T>::to_f(q[c]) * DT<T>::to_f(kb[(size_t)j * a.ldkv + c]);
    const float pr = __expf(s * scale - mx);
    l += pr;
    for (int c = 0; c < a.d; ++c) o[c] += pr * DT<T>::to_f(vb[(size_t)j * a.ldkv + c]);
  }
  T* out = reinterpret_cast<T*>(a.out) + ((size_t)b * a.S_q + i) * a.ldo + h * a.d;
  for (int c = 0; c < a.d; ++c) out[c] = DT<T>::from_f(o[c] / l);
}

void attn_simple_launch(const AttnDesc& d, cudaStream_t s) {
  const size_t total = (size_t)d.batch * d.heads * d.S_q;
  const int blocks = (int)((total + 127) / 128);
  if (d.dt == DT_F16) attn_simple_kernel<__half><<<blocks, 128, 0, s>>>(d);
  else attn_simple_kernel<__nv_bfloat16><<<blocks, 128, 0, s>>>(d);
}

template <typename T, int F, int DH> static void temporal_tiled_d(const void* qkv, void* out, int batch, int hw, int heads,
                                                                 int d, int PT, float scale, cudaStream_t s) {
  const int C = heads * d;
  const size_t smem = (size_t)F * PT * 3 * C * sizeof(T);
  const int threads = (PT * heads * F + 31) / 32 * 32;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaFuncSetAttribute(temporal_attn_tile_kernel<T, F, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_done = true;
  }
  launch_k(temporal_attn_tile_kernel<T, F, DH>, dim3(batch * (hw / PT)), dim3(threads), smem, s,
           reinterpret_cast<const T*>(qkv), reinterpret_cast<T*>(out), hw, heads, d, PT, scale);
}
template <typename T, int F> static void temporal_tiled(const void* qkv, void* out, int batch, int hw, int heads, int d,
                                                        int PT, float scale, cudaStream_t s) {
  if (F == 5 && d == 40) temporal_tiled_d<T, F, 40>(qkv, out, batch, hw, heads, d, PT, scale, s);
  else if (F == 5 && d == 80) temporal_tiled_d<T, F, 80>(qkv, out, batch, hw, heads, d, PT, scale, s);
  else if (F == 5 && d == 160) temporal_tiled_d<T, F, 160>(qkv, out, batch, hw, heads, d, PT, scale, s);
  else temporal_tiled_d<T, F, 0>(qkv, out, batch, hw, heads, d, PT, scale, s);
}

void temporal_attn_launch(int dt, const void* qkv, void* out, int batch, int frames, int hw, int heads, int d,
                          cudaStream_t s) {
  const size_t total = (size_t)batch * hw * heads;
  const int blocks = (int)((total + 127) / 128);
  const float scale = 1.0f / sqrtf((float)d);
  // wide heads (the stage-1 prior's motion modules, d = 256): d / 8 lanes per (location, head), shuffle-reduced scores
  const bool wide_on = opt(OPT_TEMPORAL_WIDE) != 0;
  // also for the UNet's head dims 40 / 80 / 160 (5 / 10 / 20 active lanes of 8 / 16 / 32): measured 0.894 -> 0.784 ms
  // per UNet forward against the tiled shared-memory kernel (2.1 -> 2.4 TB/s); OPT_TEMPORAL_WIDE_ALL = 0 restores it
  const bool wide_all = opt(OPT_TEMPORAL_WIDE_ALL) != 0;
  const bool pow2 = d == 64 || d == 128 || d == 256;
  const bool unet_d = d == 40 || d == 80 || d == 160;
  if (wide_on && frames == 5 && (pow2 || (wide_all && unet_d))) {
    const int lph = d <= 64 ? 8 : d <= 128 ? 16 : 32;
    const int wblocks = (int)((total * lph + 255) / 256);
#define RCDM_TW(T, L, V)                                                                                      \
  launch_k(temporal_attn_wide_kernel<T, 5, L, V>, dim3(wblocks), dim3(256), 0, s, reinterpret_cast<const T*>(qkv), \
           reinterpret_cast<T*>(out), batch, hw, heads, scale)
#define RCDM_TWD(T)                                  \
  switch (d) {                                       \
    case 40: RCDM_TW(T, 8, 5); break;                \
    case 64: RCDM_TW(T, 8, 8); break;                \
    case 80: RCDM_TW(T, 16, 10); break;              \
    case 128: RCDM_TW(T, 16, 16); break;             \
    case 160: RCDM_TW(T, 32, 20); break;             \
    default: RCDM_TW(T, 32, 32); break;              \
  }
    if (dt == DT_F16) {
      RCDM_TWD(__half)
    } else {
      RCDM_TWD(__nv_bfloat16)
    }
#undef RCDM_TWD
#undef RCDM_TW
    return;
  }
  // tiled kernel (coalesced through shared memory): PT = pixels per CTA, a power of two dividing hw with
  // <= 80 KB of shared memory and <= 512 threads; the one-thread-per-(pixel, head) kernel is the fallback
  const bool tiled_on = opt(OPT_TEMPORAL_TILED) != 0;
  int PT = 0;
  if (tiled_on && frames >= 1 && frames <= 5) {
    // small tiles: load / compute / store phases of a CTA do not overlap, so several CTAs per SM must
    const size_t budget = (size_t)opt(OPT_TEMPORAL_SMEM_KB) * 1024;
    const size_t per_pixel = (size_t)frames * 3 * heads * d * 2;
    for (int cand = 32; cand >= 1; cand >>= 1)
      if (hw % cand == 0 && (cand * per_pixel <= budget || cand == 1) && cand * per_pixel <= 100 * 1024 &&
          cand * heads * frames <= 512) {
        PT = cand;
        break;
      }
  }
#define RCDM_TT(T, F) temporal_tiled<T, F>(qkv, out, batch, hw, heads, d, PT, scale, s)
#define RCDM_TA(T, F)                                                                                              \
  launch_k(temporal_attn_kernel<T, F>, dim3(blocks), dim3(128), 0, s, reinterpret_cast<const T*>(qkv),         \
           reinterpret_cast<T*>(out), batch, hw, heads, d, scale)
#define RCDM_T(T, F)       \
  do {                     \
    if (PT) RCDM_TT(T, F); \
    else RCDM_TA(T, F);    \
  } while (0)
  if (dt == DT_F16) {
    switch (frames) {
      case 1: RCDM_T(__half, 1); break;
      case 2: RCDM_T(__half, 2); break;
      case 3: RCDM_T(__half, 3); break;
      case 4: RCDM_T(__half, 4); break;
      default: RCDM_T(__half, 5); break;
    }
  } else {
    switch (frames) {
      case 1: RCDM_T(__nv_bfloat16, 1); break;
      case 2: RCDM_T(__nv_bfloat16, 2); break;
      case 3: RCDM_T(__nv_bfloat16, 3); break;
      case 4: RCDM_T(__nv_bfloat16, 4); break;
      default: RCDM_T(__nv_bfloat16, 5); break;
    }
  }
#undef RCDM_T
#undef RCDM_TA
#undef RCDM_TT
}

}  // namespace rcdm

#if RCDM_ATTN_TRACE
// variant builds only: copy the timeline stamps (and the SM id of every CTA) to the host
extern "C" __attribute__((visibility("default"))) int rcdm_debug_attn_trace_read(long long* stamps, int* smids) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(stamps, rcdm::g_attn_trace, sizeof(rcdm::g_attn_trace)) != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(smids, rcdm::g_attn_smid, sizeof(rcdm::g_attn_smid)) != cudaSuccess) return 1;
  return 0;
}
#endif
