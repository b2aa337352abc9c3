// extern "C" surface of librcdm_b200.so (declared in include/rcdm.h).  No C++ exception crosses this boundary.
#include <cmath>
#include <cstring>
#include <mutex>

#include "elementwise.cuh"
#include "internal.h"
#include "prior_kernels.cuh"

using namespace rcdm;

#define API_BEGIN try {
#define API_END                                             \
  }                                                         \
  catch (const std::exception& e) { return set_err(e.what()); } \
  catch (...) { return set_err("unknown C++ exception"); }

#define CUDA_OK(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return set_err(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}

static int ensure_device_ready() {
  static std::once_flag once;
  static std::string err;
  static bool ok = false;
  std::call_once(once, [&]() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      (void)cudaGetLastError();
      err = "no CUDA device: librcdm_b200 has no CPU fallback";
      return;
    }
    ok = gemm_setup_attributes(&err) && attn_setup_attributes(&err) && gn_setup_attributes(&err) && ffn_setup_attributes(&err);
  });
  return ok ? 0 : set_err(err);
}

static inline int grid_for(size_t total, int block, int cap = 148 * 16) {
  size_t g = (total + block - 1) / block;
  return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}
static bool dt16(int dt) { return dt == RCDM_DT_F16 || dt == RCDM_DT_BF16; }

static __global__ void tap_to_f32_kernel(const void* src, int dt, float* dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = load_any(src, dt, i);
}
static void tap_to_f32(const void* src, int dt, float* dst, size_t n, cudaStream_t st) {
  tap_to_f32_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, dt, dst, n);
}

extern "C" {

const char* rcdm_version(void) { return "rcdm_b200 0.1 (sm_100a; tcgen05/TMA)"; }
const char* rcdm_last_error(void) { return g_err.c_str(); }
int rcdm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}
uint64_t rcdm_kernel_launches(void) { return g_launches.load(); }
int rcdm_debug_set_option(const char* name, int value) {
  int prev = 0;
  if (opt_set(name, value, &prev)) {
    set_err(std::string("rcdm_debug_set_option: unknown option ") + (name ? name : "(null)"));
    return -1;
  }
  return prev;
}
int rcdm_set_stream_k_min(int k_blocks) { return gemm_set_sk_min(k_blocks); }
int rcdm_set_gemm_pair(int on) { return gemm_set_pair(on); }

int rcdm_unet_create(const rcdm_unet_config* cfg, rcdm_unet** out) {
  API_BEGIN
  return unet_create(cfg, out);
  API_END
}
void rcdm_unet_destroy(rcdm_unet* h) {
  try {
    unet_destroy(h);
  } catch (...) {
  }
}
int rcdm_unet_num_weights(const rcdm_unet* h) { return h ? (int)h->slots.size() : -1; }
int rcdm_unet_weight_info(const rcdm_unet* h, int index, char* name_buf, int name_buf_len, int64_t* dims, int* ndim) {
  API_BEGIN
  if (!h || index < 0 || index >= (int)h->slots.size()) return set_err("weight index out of range");
  const Slot& s = h->slots[index];
  if (name_buf && name_buf_len > 0) {
    strncpy(name_buf, s.name.c_str(), name_buf_len - 1);
    name_buf[name_buf_len - 1] = 0;
  }
  if (dims)
    for (int i = 0; i < 4; ++i) dims[i] = s.dims[i];
  if (ndim) *ndim = s.ndim;
  return 0;
  API_END
}
int rcdm_unet_load_weight(rcdm_unet* h, const char* name, const void* data_dev, int dtype, const int64_t* dims,
                          int ndim, void* stream) {
  API_BEGIN
  return unet_load_weight(h, name, data_dev, dtype, dims, ndim, stream);
  API_END
}
int rcdm_unet_load_weights(rcdm_unet* h, int count, const char* const* names, const void* const* data_dev,
                           const int* dtypes, const int64_t* dims, const int* ndims, void* stream) {
  API_BEGIN
  return unet_load_weights(h, count, names, data_dev, dtypes, dims, ndims, stream);
  API_END
}
int rcdm_unet_weights_missing(const rcdm_unet* h) {
  if (!h) return -1;
  int n = 0;
  for (auto& s : h->slots) n += s.loaded ? 0 : 1;
  return n;
}
int rcdm_unet_prepare(rcdm_unet* h, int batch, int frames, int height, int width, int ctx_len) {
  API_BEGIN
  return unet_prepare(h, batch, frames, height, width, ctx_len);
  API_END
}
size_t rcdm_unet_workspace_bytes(const rcdm_unet* h) { return h ? h->ws_bytes : 0; }
int rcdm_unet_set_option(rcdm_unet* h, const char* name, int value) {
  if (!h || !name) return set_err("null argument");
  int* field = !strcmp(name, "simple") ? &h->simple : !strcmp(name, "autotune") ? &h->autotune :
               !strcmp(name, "ln_fold") ? &h->ln_fold : !strcmp(name, "gn_stats") ? &h->gn_stats :
               !strcmp(name, "po_fold") ? &h->po_fold : nullptr;
  if (!field) return set_err(std::string("rcdm_unet_set_option: unknown option ") + name);
  if (*field != value) {
    *field = value;
    h->planned = false;  // force a re-plan
  }
  return 0;
}
int rcdm_unet_enable_taps(rcdm_unet* h, int enable) {
  if (!h) return set_err("null handle");
  if (h->taps_enabled != (enable != 0)) {
    h->taps_enabled = enable != 0;
    h->planned = false;  // force a re-plan
  }
  return 0;
}

int rcdm_unet_forward(rcdm_unet* h, const void* sample_dev, int sample_dtype, const int64_t* timestep_dev,
                      double timestep_host, const void* ctx_dev, int ctx_dtype, void* out_dev, int out_dtype,
                      void* stream) {
  API_BEGIN
  if (!h || !sample_dev || !ctx_dev || !out_dev) return set_err("null argument");
  if (!h->planned) return set_err("rcdm_unet_forward: call rcdm_unet_prepare first");
  const int missing = rcdm_unet_weights_missing(h);
  if (missing) return set_err("rcdm_unet_forward: " + std::to_string(missing) + " state-dict entries not loaded");
  h->cur_sample = sample_dev;
  h->cur_sample_dt = sample_dtype;
  h->cur_t_dev = timestep_dev;
  h->cur_t_host = (float)timestep_host;
  h->cur_ctx = ctx_dev;
  h->cur_ctx_dt = ctx_dtype;
  h->cur_out = out_dev;
  h->cur_out_dt = out_dtype;
  return unet_run(h, true, true, reinterpret_cast<cudaStream_t>(stream));
  API_END
}

int rcdm_unet_profile(rcdm_unet* h, const void* sample_dev, int sample_dtype, double timestep_host,
                      const void* ctx_dev, int ctx_dtype, void* out_dev, int out_dtype, int reps, int max_ops,
                      float* ms_host, double* flops_host, double* bytes_host, char* kinds_host, int* dims_host,
                      int* n_ops, void* stream) {
  API_BEGIN
  if (!h || !sample_dev || !ctx_dev || !out_dev || !n_ops) return set_err("null argument");
  if (!h->planned) return set_err("rcdm_unet_profile: call rcdm_unet_prepare first");
  if (rcdm_unet_weights_missing(h)) return set_err("rcdm_unet_profile: weights not loaded");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  h->cur_sample = sample_dev;
  h->cur_sample_dt = sample_dtype;
  h->cur_t_dev = nullptr;
  h->cur_t_host = (float)timestep_host;
  h->cur_ctx = ctx_dev;
  h->cur_ctx_dt = ctx_dtype;
  h->cur_out = out_dev;
  h->cur_out_dt = out_dtype;
  if (unet_run(h, true, true, st)) return 1;  // warm-up (also finalises weights, computes the context K/V)
  const int n = (int)h->step_ops.size();
  *n_ops = n;
  if (n > max_ops) return set_err("rcdm_unet_profile: max_ops too small");
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CUDA_OK(cudaEventCreate(&e));
  std::vector<double> acc(n, 0.0);
  if (reps < 1) reps = 1;
  for (int r = 0; r < reps; ++r) {
    CUDA_OK(cudaEventRecord(ev[0], st));
    for (int i = 0; i < n; ++i) {
      h->step_ops[i](st);
      CUDA_OK(cudaEventRecord(ev[i + 1], st));
    }
    CUDA_OK(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i) {
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      acc[i] += ms;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  for (int i = 0; i < n; ++i) {
    if (ms_host) ms_host[i] = (float)(acc[i] / reps);
    if (flops_host) flops_host[i] = h->step_meta[i].flops;
    if (bytes_host) bytes_host[i] = h->step_meta[i].bytes;
    if (kinds_host) memcpy(kinds_host + (size_t)i * 16, h->step_meta[i].kind, 16);
    if (dims_host) {
      dims_host[3 * i] = h->step_meta[i].m;
      dims_host[3 * i + 1] = h->step_meta[i].n;
      dims_host[3 * i + 2] = h->step_meta[i].k;
    }
  }
  return check_launch("rcdm_unet_profile");
  API_END
}

int64_t rcdm_unet_read_tap(rcdm_unet* h, const char* name, float* out_dev, int64_t capacity, int* rows, int* channels,
                           void* stream) {
  try {
    if (!h || !name) return -(int64_t)set_err("null argument");
    auto it = h->taps.find(name);
    if (it == h->taps.end()) return -(int64_t)set_err(std::string("no such tap: ") + name);
    const TapInfo& t = it->second;
    if (rows) *rows = t.rows;
    if (channels) *channels = t.C;
    const int64_t n = (int64_t)t.rows * t.C;
    if (!out_dev) return n;
    if (n > capacity) return -(int64_t)set_err("tap buffer too small");
    tap_to_f32(h->ws + t.off, h->dt, out_dev, (size_t)n, reinterpret_cast<cudaStream_t>(stream));
    return n;
  } catch (...) {
    return -(int64_t)set_err("exception in rcdm_unet_read_tap");
  }
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// DDIM step + denoise loop
// ------------------------------------------------------------------------------------------
static void ddim_coefs(float abar_t, float abar_prev, float* c) {
  // fp32 arithmetic like the reference's 0-dim fp32 tensors (scheduling_ddim.step); c[0] is the reciprocal torch
  // forms when a tensor is divided by a CPU scalar (host IEEE division: this file's device code is --use_fast_math)
  c[0] = 1.0f / sqrtf(abar_t);
  c[1] = sqrtf(1.0f - abar_t);
  c[2] = sqrtf(abar_prev);
  c[3] = sqrtf(1.0f - abar_prev);
}

static __global__ void set_timestep_kernel(const int64_t* table, const int* step_idx, int64_t* t_cur) {
  *t_cur = table[*step_idx];
}

extern "C" {

int rcdm_ddim_cfg_step(const void* eps_dev, int eps_dtype, float* latents_f32_dev, void* latents_out_dev,
                       int latents_dtype, void* next_input_dev, int next_dtype, const void* mask_dev, int mask_dtype,
                       const void* masked_latents_dev, int masked_dtype, int clips, int frames, int height, int width,
                       int do_cfg, float guidance_scale, float alpha_bar_t, float alpha_bar_prev, void* stream) {
  API_BEGIN
  if (!eps_dev || !latents_f32_dev) return set_err("null argument");
  if (next_input_dev && (!mask_dev || !masked_latents_dev)) return set_err("next_input needs mask and masked_latents");
  if (ensure_device_ready()) return 1;
  DdimArgs a;
  memset(&a, 0, sizeof a);
  a.eps = eps_dev;
  a.eps_dt = eps_dtype;
  a.latents = latents_f32_dev;
  a.latents_out = latents_out_dev;
  a.latents_out_dt = latents_dtype;
  a.next_in = next_input_dev;
  a.next_dt = next_dtype;
  a.mask = mask_dev;
  a.mask_dt = mask_dtype;
  a.masked = masked_latents_dev;
  a.masked_dt = masked_dtype;
  a.B = clips;
  a.FHW = frames * height * width;
  a.cfg = do_cfg;
  a.guidance = guidance_scale;
  ddim_coefs(alpha_bar_t, alpha_bar_prev, a.c);
  a.round_dt = latents_dtype;
  const size_t n = (size_t)clips * 4 * a.FHW;
  ddim_cfg_step_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  g_launches++;
  return check_launch("ddim_cfg_step");
  API_END
}

int rcdm_denoise_loop(rcdm_unet* h, const void* latents_dev, int latents_dtype, const void* mask_dev, int mask_dtype,
                      const void* masked_latents_dev, int masked_dtype, const void* ctx_dev, int ctx_dtype,
                      int clips, int frames, int height, int width, int ctx_len, const int64_t* timesteps_host,
                      const float* alpha_bar_t_host, const float* alpha_bar_prev_host, int num_steps,
                      float guidance_scale, int use_graph, void* latents_out_dev, void* stream) {
  API_BEGIN
  if (!h || !latents_dev || !mask_dev || !masked_latents_dev || !ctx_dev || !latents_out_dev || !timesteps_host ||
      !alpha_bar_t_host || !alpha_bar_prev_host)
    return set_err("null argument");
  if (num_steps < 1 || clips < 1) return set_err("bad num_steps / clips");
  const int cfg = guidance_scale > 1.0f ? 1 : 0;  // RCDMs_pipeline.py:416
  const int B = cfg ? 2 * clips : clips;
  if (unet_prepare(h, B, frames, height, width, ctx_len)) return 1;
  const int missing = rcdm_unet_weights_missing(h);
  if (missing) return set_err("rcdm_denoise_loop: " + std::to_string(missing) + " state-dict entries not loaded");
  cudaStream_t caller = reinterpret_cast<cudaStream_t>(stream);
  if (!h->loop_stream) {
    CUDA_OK(cudaStreamCreateWithFlags(&h->loop_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  }
  cudaStream_t st = h->loop_stream;
  // ---- loop buffers
  const size_t FHW = (size_t)frames * height * width;
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t o_lat = 0, o_next = o_lat + al(clips * 4 * FHW * 4), o_eps = o_next + al(B * 9 * FHW * 2),
               o_tab = o_eps + al(B * 4 * FHW * 2), o_ts = o_tab + al((size_t)num_steps * 16),
               o_idx = o_ts + al((size_t)num_steps * 8), o_tcur = o_idx + 256, total = o_tcur + 256;
  if (h->loop_buf_bytes < total) {
    if (h->loop_buf) CUDA_OK(cudaFree(h->loop_buf));
    h->loop_buf = nullptr;
    CUDA_OK(cudaMalloc(&h->loop_buf, total));
    h->loop_buf_bytes = total;
    if (h->graph_exec) {
      cudaGraphExecDestroy(h->graph_exec);
      h->graph_exec = nullptr;
    }
  }
  float* lat32 = reinterpret_cast<float*>(h->loop_buf + o_lat);
  void* next_in = h->loop_buf + o_next;
  void* eps = h->loop_buf + o_eps;
  float4* table = reinterpret_cast<float4*>(h->loop_buf + o_tab);
  int64_t* ts = reinterpret_cast<int64_t*>(h->loop_buf + o_ts);
  int* step_idx = reinterpret_cast<int*>(h->loop_buf + o_idx);
  int64_t* t_cur = reinterpret_cast<int64_t*>(h->loop_buf + o_tcur);

  CUDA_OK(cudaEventRecord(h->ev_in, caller));
  CUDA_OK(cudaStreamWaitEvent(st, h->ev_in, 0));
  std::vector<float4> tab(num_steps);
  for (int i = 0; i < num_steps; ++i) {
    float c[4];
    ddim_coefs(alpha_bar_t_host[i], alpha_bar_prev_host[i], c);
    tab[i] = make_float4(c[0], c[1], c[2], c[3]);
  }
  CUDA_OK(cudaMemcpyAsync(table, tab.data(), (size_t)num_steps * 16, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(ts, timesteps_host, (size_t)num_steps * 8, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemsetAsync(step_idx, 0, 4, st));
  CUDA_OK(cudaMemsetAsync(eps, 0, (size_t)B * 4 * FHW * 2, st));
  CUDA_OK(cudaStreamSynchronize(st));  // `tab` is a host temporary
  // ---- initial state: latents -> fp32 master, first 9-channel input (identity "step": x0 = x, x_prev = x)
  DdimArgs a;
  memset(&a, 0, sizeof a);
  a.eps = eps;
  a.eps_dt = h->dt;
  a.latents = lat32;
  a.latents_out = latents_out_dev;
  a.latents_out_dt = latents_dtype;
  a.next_in = next_in;
  a.next_dt = h->dt;
  a.mask = mask_dev;
  a.mask_dt = mask_dtype;
  a.masked = masked_latents_dev;
  a.masked_dt = masked_dtype;
  a.B = clips;
  a.FHW = (int)FHW;
  a.cfg = cfg;
  a.guidance = guidance_scale;
  a.round_dt = latents_dtype;
  const size_t n = (size_t)clips * 4 * FHW;
  {
    DdimArgs init = a;
    init.c[0] = 1.f;
    init.c[1] = 0.f;
    init.c[2] = 1.f;
    init.c[3] = 0.f;
    tap_to_f32_kernel<<<grid_for(n, 256), 256, 0, st>>>(latents_dev, latents_dtype, lat32, n);
    ddim_cfg_step_kernel<<<grid_for(n, 256), 256, 0, st>>>(init);
    g_launches += 2;
  }
  // ---- step-invariant part: context cast + cross-attention K/V of all 16 spatial transformers
  h->cur_sample = next_in;
  h->cur_sample_dt = h->dt;
  h->cur_t_dev = t_cur;
  h->cur_t_host = 0.f;
  h->cur_ctx = ctx_dev;
  h->cur_ctx_dt = ctx_dtype;
  h->cur_out = eps;
  h->cur_out_dt = h->dt;
  if (unet_run(h, true, false, st)) return 1;
  // ---- one step = [t <- timesteps[i]] [UNet] [CFG + DDIM + next input] [i++]
  a.table = table;
  a.step_idx = step_idx;
  auto enqueue_step = [&]() -> int {
    set_timestep_kernel<<<1, 1, 0, st>>>(ts, step_idx, t_cur);
    if (unet_run(h, false, true, st)) return 1;
    ddim_cfg_step_kernel<<<grid_for(n, 256), 256, 0, st>>>(a);
    advance_step_kernel<<<1, 1, 0, st>>>(step_idx);
    g_launches += 3;
    return 0;
  };
  if (use_graph) {
    char key[256];
    snprintf(key, sizeof key, "%d/%d/%d/%d/%d/%d/%p/%p/%p/%d/%d/%d/%g/%p", clips, frames, height, width, ctx_len, cfg,
             mask_dev, masked_latents_dev, latents_out_dev, mask_dtype, masked_dtype, latents_dtype,
             (double)guidance_scale, (void*)h->loop_buf);
    if (!h->graph_exec || h->graph_key != key) {
      if (h->graph_exec) {
        cudaGraphExecDestroy(h->graph_exec);
        h->graph_exec = nullptr;
      }
      const uint64_t before = g_launches.load();
      CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const int rc = enqueue_step();
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      g_launches.store(before);  // captured launches are counted per replay below
      if (rc) return 1;
      if (e != cudaSuccess) return set_err(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
      e = cudaGraphInstantiate(&h->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) return set_err(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
      h->graph_key = key;
    }
    const uint64_t per_step = (uint64_t)h->step_ops.size() + 3;  // lower bound: >= 1 kernel per recorded op
    for (int i = 0; i < num_steps; ++i) CUDA_OK(cudaGraphLaunch(h->graph_exec, st));
    g_launches += per_step * num_steps;
  } else {
    for (int i = 0; i < num_steps; ++i)
      if (enqueue_step()) return 1;
  }
  CUDA_OK(cudaEventRecord(h->ev_out, st));
  CUDA_OK(cudaStreamWaitEvent(caller, h->ev_out, 0));
  return check_launch("denoise_loop");
  API_END
}

// ------------------------------------------------------------------------------------------
// single-kernel entry points
// ------------------------------------------------------------------------------------------
int rcdm_gemm(int dtype, const void* a_dev, const void* w_dev, const float* bias_dev, const void* residual_dev,
              void* out_dev, int M, int N, int K, int geglu, int tile_n, int simple, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_gemm: dtype must be f16/bf16");
  if (ensure_device_ready()) return 1;
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = M;
  d.N = N;
  d.nseg = 1;
  d.seg[0] = ASeg{SEG_PLAIN, a_dev, K, K, 0, 0, 0};
  d.w = w_dev;
  d.Ktot = K;
  d.w_rows = N;
  d.out = out_dev;
  d.ldo = geglu ? N / 2 : N;
  d.bias = bias_dev;
  d.res = residual_dev;
  d.ldr = N;
  d.geglu = geglu;
  d.force_bn = geglu ? 0 : tile_n;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (simple) {
    gemm_simple_launch(d, st);
  } else {
    GemmLaunch l;
    std::string e;
    d.sk = sk_workspace_for_stream(st, &e);
    if (!gemm_prepare(d, &l, &e)) return set_err(e);
    gemm_launch(l, st);
  }
  g_launches++;
  return check_launch("rcdm_gemm");
  API_END
}

// General form of rcdm_gemm: explicit row pitches (A rows may be strided, e.g. one token of every sample) and the
// epilogue activation of the stage-1 prior (flags: RCDM_GEMM_GEGLU | _GELU | _SILU | _SIMPLE).
int rcdm_gemm_ex(int dtype, const void* a_dev, int lda, const void* w_dev, const float* bias_dev,
                 const void* residual_dev, int ldr, void* out_dev, int ldo, int M, int N, int K, int flags,
                 void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_gemm_ex: dtype must be f16/bf16");
  if (!a_dev || !w_dev || !out_dev || M <= 0 || N <= 0 || K <= 0) return set_err("rcdm_gemm_ex: bad argument");
  const int geglu = (flags & RCDM_GEMM_GEGLU) ? 1 : 0;
  const int act = (flags & RCDM_GEMM_GELU) ? 1 : (flags & RCDM_GEMM_SILU) ? 2 : 0;
  if (geglu && (act || residual_dev)) return set_err("rcdm_gemm_ex: GEGLU excludes an activation / residual");
  if (lda < K || lda % 8) return set_err("rcdm_gemm_ex: lda must be >= K and a multiple of 8");
  if (ensure_device_ready()) return 1;
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = M;
  d.N = N;
  d.nseg = 1;
  d.seg[0] = ASeg{SEG_PLAIN, a_dev, K, lda, 0, 0, 0};
  d.w = w_dev;
  d.Ktot = K;
  d.w_rows = N;
  d.out = out_dev;
  d.ldo = ldo > 0 ? ldo : (geglu ? N / 2 : N);
  d.bias = bias_dev;
  d.res = residual_dev;
  d.ldr = ldr > 0 ? ldr : N;
  d.geglu = geglu;
  d.act = act;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (flags & RCDM_GEMM_SIMPLE) {
    gemm_simple_launch(d, st);
  } else {
    GemmLaunch l;
    std::string e;
    d.sk = sk_workspace_for_stream(st, &e);
    if (!gemm_prepare(d, &l, &e)) return set_err(e);
    gemm_launch(l, st);
  }
  g_launches++;
  return check_launch("rcdm_gemm_ex");
  API_END
}

// ---- stage-1 frame prior (prior_kernels.cuh) ----
int rcdm_masked_attn(int dtype, const void* qkv_dev, int ld, const float* key_bias_dev, int causal, void* out_dev,
                     int ldo, int batch, int heads, int S, int d, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_masked_attn: dtype must be f16/bf16");
  if (!qkv_dev || !out_dev || batch <= 0 || heads <= 0 || S <= 0) return set_err("rcdm_masked_attn: bad argument");
  if (S > 32 * MASKED_ATTN_KPL) return set_err("rcdm_masked_attn: at most 256 tokens");
  if (d <= 0 || d > 256 || d % 4) return set_err("rcdm_masked_attn: head dim must be a multiple of 4, <= 256");
  if (ld < 3 * heads * d || ld % 2 || ldo < heads * d || ldo % 2) return set_err("rcdm_masked_attn: bad row pitch");
  if (ensure_device_ready()) return 1;
  const size_t smem = masked_attn_smem_bytes(S, d);
  if (smem > 200 * 1024) return set_err("rcdm_masked_attn: S * d too large for shared memory");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // the prior's shape (d = 64, <= 112 tokens): tensor-core kernel, whole key range in registers
  const bool mma_on = opt(OPT_MASKED_ATTN_MMA) != 0;
  if (mma_on && d == MATTN_D && S <= MATTN_NT * 8 && ld % 2 == 0) {
    const float sc = 1.0f / sqrtf((float)d);
    if (dtype == DT_F16)
      masked_attn_mma_kernel<__half><<<batch * heads, MATTN_THREADS, masked_attn_mma_smem_bytes(), st>>>(
          reinterpret_cast<const __half*>(qkv_dev), ld, key_bias_dev, causal, reinterpret_cast<__half*>(out_dev), ldo, S,
          heads, sc);
    else
      masked_attn_mma_kernel<__nv_bfloat16><<<batch * heads, MATTN_THREADS, masked_attn_mma_smem_bytes(), st>>>(
          reinterpret_cast<const __nv_bfloat16*>(qkv_dev), ld, key_bias_dev, causal,
          reinterpret_cast<__nv_bfloat16*>(out_dev), ldo, S, heads, sc);
    g_launches++;
    return check_launch("rcdm_masked_attn");
  }
  const dim3 grid(batch * heads, (S + MASKED_ATTN_QCHUNK - 1) / MASKED_ATTN_QCHUNK);
  const float scale = 1.0f / sqrtf((float)d);
  if (dtype == DT_F16) {
    static bool attr = false;
    if (!attr) {
      CUDA_OK(cudaFuncSetAttribute(masked_attn_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    masked_attn_kernel<__half><<<grid, MASKED_ATTN_WARPS * 32, smem, st>>>(
        reinterpret_cast<const __half*>(qkv_dev), ld, key_bias_dev, causal, reinterpret_cast<__half*>(out_dev), ldo, S,
        heads, d, scale);
  } else {
    static bool attr = false;
    if (!attr) {
      CUDA_OK(cudaFuncSetAttribute(masked_attn_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024));
      attr = true;
    }
    masked_attn_kernel<__nv_bfloat16><<<grid, MASKED_ATTN_WARPS * 32, smem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv_dev), ld, key_bias_dev, causal,
        reinterpret_cast<__nv_bfloat16*>(out_dev), ldo, S, heads, d, scale);
  }
  g_launches++;
  return check_launch("rcdm_masked_attn");
  API_END
}

int rcdm_prior_assemble(int dtype, const void* base_dev, const void* temb_table_dev, const void* hproj_dev,
                        const void* pos_dev, void* x_dev, int batch, int S, int C, int t_row, int h_row, int n_lat,
                        const int* step_dev, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_prior_assemble: dtype must be f16/bf16");
  if (!base_dev || !temb_table_dev || !hproj_dev || !pos_dev || !x_dev) return set_err("null argument");
  if (C % 8 || n_lat <= 0 || t_row < 0 || t_row >= S || h_row < 0 || h_row >= S)
    return set_err("rcdm_prior_assemble: bad argument");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)batch * S * (C / 8);
  if (dtype == DT_F16)
    prior_assemble_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>(
        reinterpret_cast<const __half*>(base_dev), reinterpret_cast<const __half*>(temb_table_dev),
        reinterpret_cast<const __half*>(hproj_dev), reinterpret_cast<const __half*>(pos_dev),
        reinterpret_cast<__half*>(x_dev), batch, S, C, t_row, h_row, n_lat, step_dev);
  else
    prior_assemble_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(base_dev), reinterpret_cast<const __nv_bfloat16*>(temb_table_dev),
        reinterpret_cast<const __nv_bfloat16*>(hproj_dev), reinterpret_cast<const __nv_bfloat16*>(pos_dev),
        reinterpret_cast<__nv_bfloat16*>(x_dev), batch, S, C, t_row, h_row, n_lat, step_dev);
  g_launches++;
  return check_launch("rcdm_prior_assemble");
  API_END
}

int rcdm_unclip_cfg_step(int dtype, const void* pred_dev, void* latents_dev, const void* noise_table_dev,
                         const float* coef_table_dev, int n, int do_cfg, float guidance_scale, int* step_dev,
                         int advance, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_unclip_cfg_step: dtype must be f16/bf16");
  if (!pred_dev || !latents_dev || !coef_table_dev || n <= 0) return set_err("rcdm_unclip_cfg_step: bad argument");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DT_F16)
    unclip_cfg_step_kernel<__half><<<1, 1024, 0, st>>>(reinterpret_cast<const __half*>(pred_dev),
                                                       reinterpret_cast<__half*>(latents_dev),
                                                       reinterpret_cast<const __half*>(noise_table_dev), coef_table_dev,
                                                       n, do_cfg, guidance_scale, step_dev, advance);
  else
    unclip_cfg_step_kernel<__nv_bfloat16><<<1, 1024, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(pred_dev), reinterpret_cast<__nv_bfloat16*>(latents_dev),
        reinterpret_cast<const __nv_bfloat16*>(noise_table_dev), coef_table_dev, n, do_cfg, guidance_scale, step_dev,
        advance);
  g_launches++;
  return check_launch("rcdm_unclip_cfg_step");
  API_END
}

// GEMM whose epilogue also emits the per-row (sum, sum of squares) partials of its rounded output (producer side of
// the folded LayerNorm); stats_dev: float2[parts][M], *parts_out = number of column parts written
int rcdm_gemm_rowstats(int dtype, const void* a_dev, const void* w_dev, const float* bias_dev, const void* residual_dev,
                       void* out_dev, int M, int N, int K, void* stats_dev, int* parts_out, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_gemm_rowstats: dtype must be f16/bf16");
  if (!stats_dev) return set_err("null argument");
  if (ensure_device_ready()) return 1;
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = M;
  d.N = N;
  d.nseg = 1;
  d.seg[0] = ASeg{SEG_PLAIN, a_dev, K, K, 0, 0, 0};
  d.w = w_dev;
  d.Ktot = K;
  d.w_rows = N;
  d.out = out_dev;
  d.ldo = N;
  d.bias = bias_dev;
  d.res = residual_dev;
  d.ldr = N;
  d.stats_out = reinterpret_cast<float2*>(stats_dev);
  if (parts_out) *parts_out = gemm_stats_parts(N, M);
  GemmLaunch l;
  std::string e;
  d.sk = sk_workspace_for_stream(reinterpret_cast<cudaStream_t>(stream), &e);
  if (!gemm_prepare(d, &l, &e)) return set_err(e);
  gemm_launch(l, reinterpret_cast<cudaStream_t>(stream));
  g_launches++;
  return check_launch("rcdm_gemm_rowstats");
  API_END
}

// GEMM whose epilogue also accumulates the GroupNorm chunk statistics of its rounded output (see GemmParams::gn_acc):
// acc_dev = u64[M / hw][N / 10][4], zeroed by the caller; hw = rows per image.
size_t rcdm_gn_acc_bytes(int images, int channels) { return gemm_gn_acc_bytes(images, channels); }
int rcdm_gemm_gnstats(int dtype, const void* a_dev, const void* w_dev, const float* bias_dev, const void* residual_dev,
                      void* out_dev, int M, int N, int K, int hw, void* acc_dev, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_gemm_gnstats: dtype must be f16/bf16");
  if (!acc_dev) return set_err("null argument");
  if (!gemm_gn_stats_ok(M, N, hw)) return set_err("rcdm_gemm_gnstats: needs N % 160 == 0, M % 128 == 0, hw % 32 == 0");
  if (ensure_device_ready()) return 1;
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = M;
  d.N = N;
  d.nseg = 1;
  d.seg[0] = ASeg{SEG_PLAIN, a_dev, K, K, 0, 0, 0};
  d.w = w_dev;
  d.Ktot = K;
  d.w_rows = N;
  d.out = out_dev;
  d.ldo = N;
  d.bias = bias_dev;
  d.res = residual_dev;
  d.ldr = N;
  d.gn_acc = reinterpret_cast<unsigned long long*>(acc_dev);
  d.gn_hw = hw;
  GemmLaunch l;
  std::string e;
  d.sk = sk_workspace_for_stream(reinterpret_cast<cudaStream_t>(stream), &e);
  if (!gemm_prepare(d, &l, &e)) return set_err(e);
  gemm_launch(l, reinterpret_cast<cudaStream_t>(stream));
  g_launches++;
  return check_launch("rcdm_gemm_gnstats");
  API_END
}

// GroupNorm(+SiLU) of the (virtual) channel concat [x0 | x1] with the statistics taken from the accumulators the producing
// GEMMs' epilogues filled: one streaming pass, no statistics read of the tensors (x1 / acc1 may be NULL, C1 = 0).
int rcdm_groupnorm_from_stats(int dtype, const void* x0_dev, int C0, const void* acc0_dev, const void* x1_dev, int C1,
                              const void* acc1_dev, const float* gamma_dev, const float* beta_dev, void* out_dev, int rows,
                              int rows_per_stat, int hw, int groups, float eps, int silu, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_groupnorm_from_stats: dtype must be f16/bf16");
  const int C = C0 + C1;
  if (!x0_dev || !acc0_dev || !out_dev || (C1 > 0 && (!x1_dev || !acc1_dev))) return set_err("null argument");
  if (C0 % 10 || C1 % 10 || C % 8 || groups <= 0 || groups > 64 || C % groups || (C / groups) % 10 || hw <= 0 ||
      rows_per_stat % hw || rows % rows_per_stat)
    return set_err("rcdm_groupnorm_from_stats: bad shape (group width and both channel counts must be multiples of 10)");
  if (C > 3072) return set_err("rcdm_groupnorm_from_stats: at most 3072 channels");
  if (ensure_device_ready()) return 1;
  GnLaunch l;
  gn_configure_from_stats(&l, dtype, x0_dev, C0, reinterpret_cast<const unsigned long long*>(acc0_dev), x1_dev, C1,
                          reinterpret_cast<const unsigned long long*>(acc1_dev), rows, rows_per_stat, hw, groups, eps,
                          gamma_dev, beta_dev, out_dev, silu);
  gn_run(l, reinterpret_cast<cudaStream_t>(stream));
  return check_launch("rcdm_groupnorm_from_stats");
  API_END
}

size_t rcdm_linear_ln_scratch_bytes(int M, int N, int K, int frames) {
  return ((size_t)N * K * 2 + 255) / 256 * 256 + ((size_t)(frames < 1 ? 1 : frames) * N * 4 + 255) / 256 * 256 +
         (size_t)M * 8 + 256;
}

// out = (LayerNorm(x; gamma, beta, eps) [+ pe[frame(row)]]) W^T + bias, with the LayerNorm folded around the GEMM
// (weights * gamma, epilogue rstd * (acc - mean * u) + c).  geglu != 0: W / bias rows packed by rcdm_pack_geglu and the
// GEGLU gate applied (out has N/2 columns).  frame(row) = (row / rows_per_frame) % frames.
int rcdm_linear_ln(int dtype, const void* x_dev, const void* w_dev, const float* gamma_dev, const float* beta_dev,
                   const float* pe_dev, const float* bias_dev, void* out_dev, int M, int N, int K, int geglu, int frames,
                   int rows_per_frame, float eps, void* scratch_dev, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_linear_ln: dtype must be f16/bf16");
  if (!x_dev || !w_dev || !gamma_dev || !beta_dev || !out_dev || !scratch_dev) return set_err("null argument");
  if (frames < 1) frames = 1;
  if (frames > 5) return set_err("rcdm_linear_ln: at most 5 frames");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* sc = reinterpret_cast<unsigned char*>(scratch_dev);
  void* wf = sc;
  size_t off = ((size_t)N * K * 2 + 255) / 256 * 256;
  float* c = reinterpret_cast<float*>(sc + off);
  off += ((size_t)frames * N * 4 + 255) / 256 * 256;
  float2* stats = reinterpret_cast<float2*>(sc + off);
  const int fb = (N * 32 + 255) / 256, rb = (M * 32 + 255) / 256;
  if (dtype == DT_F16) {
    fold_ln_kernel<__half><<<fb, 256, 0, st>>>(reinterpret_cast<const __half*>(w_dev), reinterpret_cast<__half*>(wf),
                                               gamma_dev, beta_dev, pe_dev, bias_dev, c, N, K, frames);
    rowstats_kernel<__half><<<rb, 256, 0, st>>>(reinterpret_cast<const __half*>(x_dev), stats, M, K);
  } else {
    fold_ln_kernel<__nv_bfloat16><<<fb, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(w_dev),
                                                      reinterpret_cast<__nv_bfloat16*>(wf), gamma_dev, beta_dev, pe_dev,
                                                      bias_dev, c, N, K, frames);
    rowstats_kernel<__nv_bfloat16><<<rb, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x_dev), stats, M, K);
  }
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = M;
  d.N = N;
  d.nseg = 1;
  d.seg[0] = ASeg{SEG_PLAIN, x_dev, K, K, 0, 0, 0};
  d.w = wf;
  d.Ktot = K;
  d.w_rows = N;
  d.out = out_dev;
  d.ldo = geglu ? N / 2 : N;
  d.geglu = geglu;
  d.stats_in = stats;
  d.stats_parts = 1;
  d.ln_c = c;
  d.ln_frames = frames;
  d.ln_rows_per_frame = rows_per_frame;
  d.ln_eps = eps;
  GemmLaunch l;
  std::string e;
  d.sk = sk_workspace_for_stream(st, &e);
  if (!gemm_prepare(d, &l, &e)) return set_err(e);
  gemm_launch(l, st);
  g_launches += 3;
  return check_launch("rcdm_linear_ln");
  API_END
}

// The two halves of rcdm_linear_ln as separate calls, for a host that folds its weights once at load time and chains
// GEMMs (the stage-1 prior's Python host): rcdm_fold_ln = the load-time half (centred gamma-scaled weights + constant
// vector), rcdm_gemm_ln = rcdm_gemm_ex that can consume row statistics (folded LayerNorm in front of it) and / or emit
// the statistics of its own rounded output for the next folded LayerNorm.
int rcdm_fold_ln(int dtype, const void* w_dev, const float* gamma_dev, const float* beta_dev, const float* pe_dev,
                 const float* bias_dev, void* wf_out_dev, float* c_out_dev, int N, int K, int frames, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_fold_ln: dtype must be f16/bf16");
  if (!w_dev || !gamma_dev || !beta_dev || !wf_out_dev || !c_out_dev || N <= 0 || K <= 0) return set_err("rcdm_fold_ln: bad argument");
  if (frames < 1) frames = 1;
  if (frames > 5) return set_err("rcdm_fold_ln: at most 5 frames");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int fb = (N * 32 + 255) / 256;
  if (dtype == DT_F16)
    fold_ln_kernel<__half><<<fb, 256, 0, st>>>(reinterpret_cast<const __half*>(w_dev), reinterpret_cast<__half*>(wf_out_dev),
                                               gamma_dev, beta_dev, pe_dev, bias_dev, c_out_dev, N, K, frames);
  else
    fold_ln_kernel<__nv_bfloat16><<<fb, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(w_dev),
                                                      reinterpret_cast<__nv_bfloat16*>(wf_out_dev), gamma_dev, beta_dev,
                                                      pe_dev, bias_dev, c_out_dev, N, K, frames);
  g_launches++;
  return check_launch("rcdm_fold_ln");
  API_END
}

int rcdm_gemm_stats_parts(int M, int N) { return gemm_stats_parts(N, M); }

int rcdm_rowstats(int dtype, const void* x_dev, void* stats_dev, int M, int K, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_rowstats: dtype must be f16/bf16");
  if (!x_dev || !stats_dev || M <= 0 || K <= 0) return set_err("rcdm_rowstats: bad argument");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rb = (M * 32 + 255) / 256;
  if (dtype == DT_F16)
    rowstats_kernel<__half><<<rb, 256, 0, st>>>(reinterpret_cast<const __half*>(x_dev), reinterpret_cast<float2*>(stats_dev), M, K);
  else
    rowstats_kernel<__nv_bfloat16><<<rb, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x_dev),
                                                       reinterpret_cast<float2*>(stats_dev), M, K);
  g_launches++;
  return check_launch("rcdm_rowstats");
  API_END
}

int rcdm_gemm_ln(int dtype, const void* a_dev, int lda, const void* w_dev, const float* vec_dev, const void* residual_dev,
                 void* out_dev, int M, int N, int K, int flags, const void* stats_in_dev, int parts_in, int frames,
                 int rows_per_frame, float eps, void* stats_out_dev, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_gemm_ln: dtype must be f16/bf16");
  if (!a_dev || !w_dev || !out_dev || M <= 0 || N <= 0 || K <= 0) return set_err("rcdm_gemm_ln: bad argument");
  const int geglu = (flags & RCDM_GEMM_GEGLU) ? 1 : 0;
  const int act = (flags & RCDM_GEMM_GELU) ? 1 : (flags & RCDM_GEMM_SILU) ? 2 : 0;
  if (flags & RCDM_GEMM_SIMPLE) return set_err("rcdm_gemm_ln: no CUDA-core form (use rcdm_layernorm + rcdm_gemm_ex)");
  if (geglu && (act || residual_dev || stats_out_dev)) return set_err("rcdm_gemm_ln: GEGLU excludes an activation / residual / statistics");
  if (lda < K || lda % 8) return set_err("rcdm_gemm_ln: lda must be >= K and a multiple of 8");
  if (stats_in_dev && (!vec_dev || parts_in <= 0)) return set_err("rcdm_gemm_ln: folded LayerNorm needs the c vector and parts_in > 0");
  if (stats_in_dev && lda != K) return set_err("rcdm_gemm_ln: folded LayerNorm needs dense A rows");
  if (frames < 1) frames = 1;
  if (frames > 5) return set_err("rcdm_gemm_ln: at most 5 frames");
  if (ensure_device_ready()) return 1;
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = M;
  d.N = N;
  d.nseg = 1;
  d.seg[0] = ASeg{SEG_PLAIN, a_dev, K, lda, 0, 0, 0};
  d.w = w_dev;
  d.Ktot = K;
  d.w_rows = N;
  d.out = out_dev;
  d.ldo = geglu ? N / 2 : N;
  d.res = residual_dev;
  d.ldr = N;
  d.geglu = geglu;
  d.act = act;
  if (stats_in_dev) {
    d.stats_in = reinterpret_cast<const float2*>(stats_in_dev);
    d.stats_parts = parts_in;
    d.ln_c = vec_dev;
    d.ln_frames = frames;
    d.ln_rows_per_frame = rows_per_frame;
    d.ln_eps = eps;
  } else {
    d.bias = vec_dev;
  }
  d.stats_out = reinterpret_cast<float2*>(stats_out_dev);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GemmLaunch l;
  std::string e;
  d.sk = sk_workspace_for_stream(st, &e);
  if (!gemm_prepare(d, &l, &e)) return set_err(e);
  gemm_launch(l, st);
  g_launches++;
  return check_launch("rcdm_gemm_ln");
  API_END
}

// proj_out folded over ff.net.2 (fold_proj_kernel, unet.cu) and the two-segment GEMM that runs on the folded weights, for a
// host that chains GEMMs itself (stage-1 prior: motion_module.py:170-180,243).
int rcdm_fold_proj(int dtype, const void* wp_dev, const void* w2_dev, const float* b2_dev, const float* bp_dev,
                   void* wf_out_dev, float* cf_out_dev, int C, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_fold_proj: dtype must be f16/bf16");
  if (!wp_dev || !w2_dev || !b2_dev || !bp_dev || !wf_out_dev || !cf_out_dev || C <= 0) return set_err("rcdm_fold_proj: bad argument");
  if (ensure_device_ready()) return 1;
  fold_proj_launch(dtype, wp_dev, w2_dev, b2_dev, bp_dev, wf_out_dev, cf_out_dev, C, reinterpret_cast<cudaStream_t>(stream));
  g_launches++;
  return check_launch("rcdm_fold_proj");
  API_END
}

int rcdm_gemm_cat(int dtype, const void* a0_dev, int K0, const void* a1_dev, int K1, const void* w_dev, const float* bias_dev,
                  const void* residual_dev, void* out_dev, int M, int N, void* stats_out_dev, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_gemm_cat: dtype must be f16/bf16");
  if (!a0_dev || !a1_dev || !w_dev || !out_dev || M <= 0 || N <= 0 || K0 <= 0 || K1 <= 0) return set_err("rcdm_gemm_cat: bad argument");
  if (K0 % 64 || K1 % 64) return set_err("rcdm_gemm_cat: K0 and K1 must be multiples of 64");
  if (ensure_device_ready()) return 1;
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = M;
  d.N = N;
  d.nseg = 2;
  d.seg[0] = ASeg{SEG_PLAIN, a0_dev, K0, K0, 0, 0, 0};
  d.seg[1] = ASeg{SEG_PLAIN, a1_dev, K1, K1, 0, 0, 0};
  d.w = w_dev;
  d.Ktot = K0 + K1;
  d.w_rows = N;
  d.out = out_dev;
  d.ldo = N;
  d.bias = bias_dev;
  d.res = residual_dev;
  d.ldr = N;
  d.stats_out = reinterpret_cast<float2*>(stats_out_dev);
  if (stats_out_dev) d.force_bn = gemm_plain_bn(M, N);  // the consumer sums gemm_stats_parts(N, M) parts: same tile width
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GemmLaunch l;
  std::string e;
  d.sk = sk_workspace_for_stream(st, &e);
  if (!gemm_prepare(d, &l, &e)) return set_err(e);
  gemm_launch(l, st);
  g_launches++;
  return check_launch("rcdm_gemm_cat");
  API_END
}

// Fused GEGLU feed-forward of the C = 320 transformer blocks (ffn_fused.cuh): out = y + GEGLU(LayerNorm(y) W1^T + b1) W2^T + b2.
// w1 [2560, 320] / bias1 [2560] in the reference layout (h rows, then gate rows), w2 [320, 1280].  The weight folding /
// packing and the row statistics of y (done once at load time / by the producing GEMM inside the UNet plan) run here per call.
size_t rcdm_ffn_geglu_scratch_bytes(int M) {
  const size_t wbytes = (size_t)2 * FfnCfg::J * FfnCfg::C * 2;
  return 2 * wbytes + 2 * ((size_t)2 * FfnCfg::J * 4 + 256) + (size_t)(M > 0 ? M : 0) * 8 + 1024;
}
int rcdm_ffn_geglu_ln(int dtype, const void* y_dev, const void* w1_dev, const float* gamma_dev, const float* beta_dev,
                      const float* bias1_dev, const void* w2_dev, const float* bias2_dev, void* out_dev, int M, float eps,
                      void* scratch_dev, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_ffn_geglu_ln: dtype must be f16/bf16");
  if (!y_dev || !w1_dev || !gamma_dev || !beta_dev || !bias1_dev || !w2_dev || !out_dev || !scratch_dev || M <= 0)
    return set_err("rcdm_ffn_geglu_ln: bad argument");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  constexpr int N = 2 * FfnCfg::J, K = FfnCfg::C, BN = FFN_FUSED_GEGLU_BN;
  unsigned char* sc = reinterpret_cast<unsigned char*>(scratch_dev);
  const size_t wbytes = (size_t)N * K * 2;
  void* wp = sc;
  void* wf = sc + wbytes;
  float* bp = reinterpret_cast<float*>(sc + 2 * wbytes);
  float* c = reinterpret_cast<float*>(sc + 2 * wbytes + ((size_t)N * 4 + 256));
  float2* stats = reinterpret_cast<float2*>(sc + 2 * wbytes + 2 * ((size_t)N * 4 + 256));
  const int fb = (N * 32 + 255) / 256, rb = (M * 32 + 255) / 256;
  pack_vec_kernel<<<(N + 255) / 256, 256, 0, st>>>(bias1_dev, DT_F32, bp, N, 0, BN, 0);
  if (dtype == DT_F16) {
    pack_weight_kernel<__half><<<grid_for((size_t)N * K, 256), 256, 0, st>>>(w1_dev, dtype, reinterpret_cast<__half*>(wp), N, K, K,
                                                                             0, 0, 0, 0, BN);
    fold_ln_kernel<__half><<<fb, 256, 0, st>>>(reinterpret_cast<const __half*>(wp), reinterpret_cast<__half*>(wf), gamma_dev,
                                               beta_dev, nullptr, bp, c, N, K, 1);
    rowstats_kernel<__half><<<rb, 256, 0, st>>>(reinterpret_cast<const __half*>(y_dev), stats, M, K);
  } else {
    pack_weight_kernel<__nv_bfloat16><<<grid_for((size_t)N * K, 256), 256, 0, st>>>(
        w1_dev, dtype, reinterpret_cast<__nv_bfloat16*>(wp), N, K, K, 0, 0, 0, 0, BN);
    fold_ln_kernel<__nv_bfloat16><<<fb, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(wp),
                                                      reinterpret_cast<__nv_bfloat16*>(wf), gamma_dev, beta_dev, nullptr, bp, c,
                                                      N, K, 1);
    rowstats_kernel<__nv_bfloat16><<<rb, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(y_dev), stats, M, K);
  }
  FfnDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = M;
  d.y = y_dev;
  d.stats_in = stats;
  d.stats_parts = 1;
  d.ln_eps = eps;
  d.w1f = wf;
  d.c1 = c;
  d.w2 = w2_dev;
  d.bias2 = bias2_dev;
  d.out = out_dev;
  d.pair = opt(OPT_FFN_FUSED) >= 2 ? 1 : 0;
  FfnLaunch l;
  std::string e;
  if (!ffn_prepare(d, &l, &e)) return set_err(e);
  ffn_launch(l, st);
  g_launches += 5;
  return check_launch("rcdm_ffn_geglu_ln");
  API_END
}

int rcdm_pack_geglu(int dtype, const void* w_dev, const float* bias_dev, void* w_out_dev, float* bias_out_dev, int N,
                    int K, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("dtype must be f16/bf16");
  const int GEGLU_BN = geglu_bn(N);
  if (N % GEGLU_BN) return set_err("GEGLU pack: N must be a multiple of 128 (or of 160)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DT_F16)
    pack_weight_kernel<__half><<<grid_for((size_t)N * K, 256), 256, 0, st>>>(
        w_dev, dtype, reinterpret_cast<__half*>(w_out_dev), N, K, K, 0, 0, 0, 0, GEGLU_BN);
  else
    pack_weight_kernel<__nv_bfloat16><<<grid_for((size_t)N * K, 256), 256, 0, st>>>(
        w_dev, dtype, reinterpret_cast<__nv_bfloat16*>(w_out_dev), N, K, K, 0, 0, 0, 0, GEGLU_BN);
  if (bias_dev && bias_out_dev)
    pack_vec_kernel<<<(N + 255) / 256, 256, 0, st>>>(bias_dev, DT_F32, bias_out_dev, N, 0, GEGLU_BN, 0);
  g_launches += 2;
  return check_launch("rcdm_pack_geglu");
  API_END
}

int rcdm_pack_conv3x3(int dtype, const void* w_dev, void* w_out_dev, int cout, int cin, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("dtype must be f16/bf16");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)cout * cin * 9;
  if (dtype == DT_F16)
    pack_weight_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>(w_dev, dtype, reinterpret_cast<__half*>(w_out_dev),
                                                                     cout, 9 * cin, 9 * cin, 0, 0, 1, cin, 0);
  else
    pack_weight_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(
        w_dev, dtype, reinterpret_cast<__nv_bfloat16*>(w_out_dev), cout, 9 * cin, 9 * cin, 0, 0, 1, cin, 0);
  g_launches++;
  return check_launch("rcdm_pack_conv3x3");
  API_END
}

int rcdm_conv3x3(int dtype, const void* x_dev, const void* w_packed_dev, const float* bias_dev,
                 const void* residual_dev, void* out_dev, int n, int h, int w, int cin, int cout, int stride,
                 int simple, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_conv3x3: dtype must be f16/bf16");
  if (stride != 1 && stride != 2 && stride != -2) return set_err("stride must be 1, 2 or -2 (stride 2, padding (0,1,0,1))");
  if (ensure_device_ready()) return 1;
  const int seg_mode = stride == 1 ? SEG_CONV3 : stride == 2 ? SEG_CONV3S2 : SEG_CONV3S2A;
  if (stride < 0) stride = 2;
  const int Ho = h / stride, Wo = w / stride;
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.dt = dtype;
  d.M = n * Ho * Wo;
  d.N = cout;
  d.nseg = 1;
  d.seg[0] = ASeg{seg_mode, x_dev, cin, cin, h, w, n};
  d.w = w_packed_dev;
  d.Ktot = 9 * cin;
  d.w_rows = cout;
  d.Ho = Ho;
  d.Wo = Wo;
  d.NI = n;
  d.out = out_dev;
  d.ldo = cout;
  d.bias = bias_dev;
  d.res = residual_dev;
  d.ldr = cout;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (simple) {
    gemm_simple_launch(d, st);
  } else {
    GemmLaunch l;
    std::string e;
    d.sk = sk_workspace_for_stream(st, &e);
    if (!gemm_prepare(d, &l, &e)) return set_err(e);
    gemm_launch(l, st);
  }
  g_launches++;
  return check_launch("rcdm_conv3x3");
  API_END
}

// Upsample (nearest 2x) + conv3x3 / pad 1 without the 4x tensor: fold the weights per output parity class, then four
// tensor-core launches of a 2x2 conv on the original activation (K = 4 cin).  wf_dev: scratch / cache for the folded
// weights, rcdm_upsample_conv3x3_weight_bytes(cout, cin) bytes; fold != 0 (re)computes it from w_packed_dev.
size_t rcdm_upsample_conv3x3_weight_bytes(int cout, int cin) { return (size_t)16 * cout * cin * 2; }
int rcdm_upsample_conv3x3(int dtype, const void* x_dev, const void* w_packed_dev, const float* bias_dev, void* out_dev,
                          int n, int h, int w, int cin, int cout, void* wf_dev, int fold, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_upsample_conv3x3: dtype must be f16/bf16");
  if (!x_dev || !out_dev || !wf_dev || (fold && !w_packed_dev)) return set_err("null argument");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (fold) {
    const size_t total = (size_t)16 * cout * cin;
    if (dtype == DT_F16)
      fold_upsample_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __half*>(w_packed_dev),
                                                                        reinterpret_cast<__half*>(wf_dev), cout, cin);
    else
      fold_upsample_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(w_packed_dev), reinterpret_cast<__nv_bfloat16*>(wf_dev), cout, cin);
    g_launches++;
  }
  for (int cls = 0; cls < 4; ++cls) {
    GemmDesc d;
    memset(&d, 0, sizeof d);
    d.dt = dtype;
    d.M = n * h * w;
    d.N = cout;
    d.nseg = 1;
    d.seg[0] = ASeg{SEG_UP2, x_dev, cin, cin, h, w, n};
    d.w = reinterpret_cast<const char*>(wf_dev) + (size_t)cls * cout * 4 * cin * 2;
    d.Ktot = 4 * cin;
    d.w_rows = cout;
    d.Ho = h;
    d.Wo = w;
    d.NI = n;
    d.out = out_dev;
    d.ldo = cout;
    d.bias = bias_dev;
    d.up_py = cls >> 1;
    d.up_px = cls & 1;
    GemmLaunch l;
    std::string e;
    d.sk = sk_workspace_for_stream(st, &e);
    if (!gemm_prepare(d, &l, &e)) return set_err(e);
    gemm_launch(l, st);
    g_launches++;
  }
  return check_launch("rcdm_upsample_conv3x3");
  API_END
}

// ---- the pieces AutoencoderKL needs beyond the UNet's kernels (RCDMs_pipeline.py:274-287,429-431) ----
int rcdm_conv3x3_small(int dtype, const void* x_dev, const void* w_packed_dev, const float* bias_dev, void* out_dev, int n,
                       int h, int w, int cin, int cout, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_conv3x3_small: dtype must be f16/bf16");
  if (!x_dev || !w_packed_dev || !out_dev) return set_err("null argument");
  if (cout % 8 || cin < 1 || cin > 16) return set_err("rcdm_conv3x3_small: cout % 8 == 0 and cin <= 16");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)n * h * w * (cout / 8);
  if (dtype == DT_F16)
    conv3x3_small_kernel<__half><<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(
        reinterpret_cast<const __half*>(x_dev), reinterpret_cast<const __half*>(w_packed_dev), bias_dev,
        reinterpret_cast<__half*>(out_dev), n, h, w, cin, cout);
  else
    conv3x3_small_kernel<__nv_bfloat16><<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_dev), reinterpret_cast<const __nv_bfloat16*>(w_packed_dev), bias_dev,
        reinterpret_cast<__nv_bfloat16*>(out_dev), n, h, w, cin, cout);
  g_launches++;
  return check_launch("rcdm_conv3x3_small");
  API_END
}

int rcdm_linear_small(int dtype, const void* x_dev, const void* w_dev, const float* bias_dev, void* out_dev, int64_t M, int N,
                      int K, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_linear_small: dtype must be f16/bf16");
  if (!x_dev || !w_dev || !out_dev || M <= 0 || N < 1 || N > 16 || K < 1 || K > 16) return set_err("rcdm_linear_small: N, K <= 16");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DT_F16)
    linear_small_kernel<__half><<<grid_for((size_t)M, 256, 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const __half*>(x_dev), reinterpret_cast<const __half*>(w_dev), bias_dev,
        reinterpret_cast<__half*>(out_dev), (size_t)M, N, K);
  else
    linear_small_kernel<__nv_bfloat16><<<grid_for((size_t)M, 256, 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_dev), reinterpret_cast<const __nv_bfloat16*>(w_dev), bias_dev,
        reinterpret_cast<__nv_bfloat16*>(out_dev), (size_t)M, N, K);
  g_launches++;
  return check_launch("rcdm_linear_small");
  API_END
}

int rcdm_upsample2x(int dtype, const void* x_dev, void* out_dev, int n, int h, int w, int c, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_upsample2x: dtype must be f16/bf16");
  if (!x_dev || !out_dev || c % 8) return set_err("rcdm_upsample2x: bad argument (c % 8 == 0)");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = (size_t)n * 4 * h * w * (c / 8);
  if (dtype == DT_F16)
    upsample2x_kernel<__half><<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(reinterpret_cast<const __half*>(x_dev),
                                                                              reinterpret_cast<__half*>(out_dev), n, h, w, c);
  else
    upsample2x_kernel<__nv_bfloat16><<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_dev), reinterpret_cast<__nv_bfloat16*>(out_dev), n, h, w, c);
  g_launches++;
  return check_launch("rcdm_upsample2x");
  API_END
}

int rcdm_softmax_rows(int dtype, void* x_dev, int rows, int cols, int ld, float scale, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_softmax_rows: dtype must be f16/bf16");
  if (!x_dev || rows <= 0 || cols <= 0 || cols % 8 || ld % 8 || ld < cols) return set_err("rcdm_softmax_rows: bad argument");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int threads = cols >= 2048 ? 256 : 128;
  if (dtype == DT_F16) softmax_rows_kernel<__half><<<rows, threads, 0, st>>>(reinterpret_cast<__half*>(x_dev), cols, ld, scale);
  else softmax_rows_kernel<__nv_bfloat16><<<rows, threads, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(x_dev), cols, ld, scale);
  g_launches++;
  return check_launch("rcdm_softmax_rows");
  API_END
}

size_t rcdm_groupnorm_scratch_bytes(int rows, int rows_per_stat, int groups) {
  if (rows_per_stat <= 0) return 0;
  return gn_scratch_bytes(rows / rows_per_stat, groups);
}

int rcdm_groupnorm(int dtype, const void* x_dev, const float* gamma_dev, const float* beta_dev, void* out_dev,
                   int rows, int channels, int groups, int rows_per_stat, float eps, int silu, void* scratch_dev,
                   void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_groupnorm: dtype must be f16/bf16");
  if (channels % 8 || channels % groups || rows % rows_per_stat) return set_err("rcdm_groupnorm: bad shape");
  if (rows / rows_per_stat > 16384) return set_err("rcdm_groupnorm: too many statistic batches");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUDA_OK(cudaMemsetAsync(scratch_dev, 0, 65536, st));
  GnLaunch l;
  gn_configure(&l, dtype, x_dev, channels, nullptr, 0, rows, rows_per_stat, groups, eps, gamma_dev, beta_dev, out_dev,
               silu, scratch_dev);
  gn_run(l, st);
  return check_launch("rcdm_groupnorm");
  API_END
}

int rcdm_layernorm(int dtype, const void* x_dev, const float* gamma_dev, const float* beta_dev, void* out_dev,
                   int rows, int channels, float eps, const float* pe_dev, int rows_per_frame, int frames,
                   void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_layernorm: dtype must be f16/bf16");
  if (ensure_device_ready()) return 1;
  if (!ln_run(dtype, x_dev, out_dev, gamma_dev, beta_dev, rows, channels, eps, pe_dev, rows_per_frame > 0 ? rows_per_frame : 1,
              frames > 0 ? frames : 1, reinterpret_cast<cudaStream_t>(stream)))
    return set_err("rcdm_layernorm: unsupported channel count");
  return check_launch("rcdm_layernorm");
  API_END
}

int rcdm_flash_attn(int dtype, const void* q_dev, int ldq, const void* k_dev, const void* v_dev, int ldkv,
                    void* out_dev, int ldo, int batch, int heads, int S_q, int S_kv, int d, int simple, void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_flash_attn: dtype must be f16/bf16");
  if (ensure_device_ready()) return 1;
  AttnDesc a;
  memset(&a, 0, sizeof a);
  a.dt = dtype;
  a.q = q_dev;
  a.ldq = ldq;
  a.k = k_dev;
  a.v = v_dev;
  a.ldkv = ldkv;
  a.S_q = S_q;
  a.S_kv = S_kv;
  a.heads = heads;
  a.d = d;
  a.batch = batch;
  a.out = out_dev;
  a.ldo = ldo;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (simple) {
    if (d > 160) return set_err("head dim > 160");
    attn_simple_launch(a, st);
  } else {
    AttnLaunch l;
    std::string e;
    if (!attn_prepare(a, &l, &e)) return set_err(e);
    attn_launch(l, st);
  }
  g_launches++;
  return check_launch("rcdm_flash_attn");
  API_END
}

int rcdm_temporal_attn(int dtype, const void* qkv_dev, void* out_dev, int batch, int frames, int hw, int heads, int d,
                       void* stream) {
  API_BEGIN
  if (!dt16(dtype)) return set_err("rcdm_temporal_attn: dtype must be f16/bf16");
  if (frames < 1 || frames > 5 || d % 8) return set_err("rcdm_temporal_attn: frames must be 1..5 and d % 8 == 0");
  if (ensure_device_ready()) return 1;
  temporal_attn_launch(dtype, qkv_dev, out_dev, batch, frames, hw, heads, d, reinterpret_cast<cudaStream_t>(stream));
  g_launches++;
  return check_launch("rcdm_temporal_attn");
  API_END
}

}  // extern "C"
