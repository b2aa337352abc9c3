// Host-side launch descriptors shared by the translation units of librcdm_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>

#include "attention.cuh"
#include "gemm_tcgen05.cuh"
#include "ffn_fused.cuh"

namespace rcdm {

// ---- debug / experiment switches -------------------------------------------------------------------------------
// The product library reads NO environment variables: every switch has a compiled-in default (the measured-best
// setting) and can only be changed explicitly through the ABI (rcdm_debug_set_option), which tests and the profiling
// scripts use to bisect a parity failure or to reproduce a rejected variant.
enum Opt : int {
  OPT_PDL = 0,            // 0 off (default: measured slower inside CUDA graphs) | 1 programmatic dependent launch
  OPT_SK_MIN,             // k-blocks a launch must save before stream-K is used (default 24; 0 = never)
  OPT_GEMM_PAIR,          // 0 never | 1 heuristic (default) | 2 whenever there are >= 2 M tiles
  OPT_MASKED_ATTN_MMA,    // stage-1 prior: tensor-core masked attention (default 1)
  OPT_TEMPORAL_WIDE,      // warp-slice temporal attention for d = 64/128/256 (default 1)
  OPT_TEMPORAL_WIDE_ALL,  // ... also for the UNet's d = 40/80/160 (default 1)
  OPT_TEMPORAL_TILED,     // tiled shared-memory temporal attention for the remaining head dims (default 1)
  OPT_TEMPORAL_SMEM_KB,   // its shared-memory budget per CTA (default 40)
  OPT_GN_FUSED,           // single-launch GroupNorm with a grid barrier (default 1; 0 = two-kernel path)
  OPT_LN_WIDE,            // CTA-per-row LayerNorm for rows wider than 1280 (default 1)
  OPT_GN_STATS,           // GroupNorm statistics from the producing GEMM's epilogue + streaming apply (default 1)
  OPT_ATTN_SHORT_KV,      // attention over <= 112 keys (the UNet's cross-attention) on the register-resident mma.sync kernel
                          // (default 1: 217 vs 292 us at 80 x 4096 queries x 91 keys, profiles/r02_xattn_bench_v3.txt; 0 = tcgen05 flash
                          // kernel for every key range)
  OPT_FFN_FUSED,          // fused GEGLU feed-forward kernel for the C = 320 transformer blocks (default 0: measured slower
                          // than the two GEMMs; set BEFORE rcdm_unet_create - it decides the weight packing)
  OPT_COUNT
};
int opt(int id);
int opt_set(const char* name, int value, int* previous);  // 0 = ok, 1 = unknown name

// ---- kernel launch with programmatic stream serialization (PDL); off by default (OPT_PDL) ----
bool pdl_enabled();
int pdl_mode();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
// Cooperative launch: for the kernels whose CTAs synchronise through global memory (the GroupNorm grid barrier, the
// stream-K fix-up): the driver GUARANTEES that all CTAs are co-resident or refuses the launch (an error, never a spin
// that ends in a trap when another stream / process holds SMs).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_coop(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- TMA tensor-map encoding (driver entry point resolved at run time; no link-time libcuda dependency) ----
// dims/box innermost-first; strides_bytes has rank-1 entries (dimension 0 is contiguous). 16-bit elements.
bool encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, bool swizzle128, std::string* err);
bool encode_tmap_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes /*0 | 32 | 64 | 128*/, std::string* err);

// ---- GEMM / implicit-GEMM conv -------------------------------------------------------------------------
struct ASeg {
  int mode;         // SEG_PLAIN / SEG_CONV3 / SEG_CONV3S2 / SEG_CONV3S2A / SEG_UP2
  const void* ptr;  // plain: [M, ld]; conv: NHWC activation [NI, H, W, C]
  int C;            // K of this segment per tap
  int ld;           // plain: row pitch in elements
  int H, W, NI;     // conv: INPUT spatial dims and image count
};
// stream-K partial-tile workspace (gemm_host.cu): owned by whoever serialises the launches that use it
struct SkWorkspace {
  float* ws = nullptr;        // [slots][128 * 192] fp32
  unsigned* flags = nullptr;  // [slots], zero at rest
  int slots = 0;
  int device = -1;
};
bool sk_workspace_alloc(SkWorkspace* w, std::string* err);  // on the current device; no-op when already allocated
void sk_workspace_free(SkWorkspace* w);
// per-(device, stream) workspace for the stand-alone entry points (lives for the process)
const SkWorkspace* sk_workspace_for_stream(cudaStream_t s, std::string* err);

struct GemmDesc {
  int dt;  // DT_F16 / DT_BF16
  int M, N;
  int nseg;
  ASeg seg[3];
  const void* w;  // [w_rows, Ktot] row-major, K order = segments in sequence (conv: tap-major, channel-minor)
  int Ktot, w_rows;
  int Ho, Wo, NI;  // conv OUTPUT grid (M = NI*Ho*Wo); ignored for plain-only GEMMs
  void* out;
  int ldo;
  const float* bias;
  const void* res;
  int ldr;
  int geglu;
  int act;       // 1 = erf GELU, 2 = SiLU on (acc + bias), before the residual add
  int force_bn;  // 0 = heuristic
  int no_sk;     // 1 = never use the stream-K decomposition for this launch
  const SkWorkspace* sk;  // stream-K workspace of the launching stream / handle; nullptr = no stream-K
  int no_pair;   // 1 = never use the CTA-pair (cta_group::2) kernel for this launch
  int force_pair;  // 1 = always use it (when there are >= 2 M tiles); set by the plan-time autotuner
  // folded LayerNorm (see GemmParams): producer side / consumer side
  float2* stats_out;
  const float2* stats_in;
  int stats_parts, ln_frames, ln_rows_per_frame;
  const float* ln_c;
  float ln_eps;
  // GroupNorm statistics of the output from the epilogue (see GemmParams::gn_acc); requires N % 160 == 0, M % 128 == 0,
  // gn_hw % 32 == 0 (gemm_gn_stats_ok)
  unsigned long long* gn_acc;
  int gn_hw;
  // SEG_UP2 (upsample folded into the conv): parity class of this launch; `out` is the FULL-resolution tensor
  // [NI, 2 Ho, 2 Wo, ldo] and the class's pixels are written through a strided 4-D map
  int up_py, up_px;
};
bool gemm_gn_stats_ok(int M, int N, int hw);  // can a GEMM / conv with this output shape emit GroupNorm statistics?
inline size_t gemm_gn_acc_bytes(int n_img, int N) { return (size_t)n_img * (N / GN_CHUNK) * 4 * sizeof(unsigned long long); }
struct GemmLaunch {
  GemmMaps maps;
  GemmParams p;
  dim3 grid;
  int bn, dt;
  int pair;  // 1 = CTA-pair kernel (cluster of 2, 256-row tiles)
  int gn;    // 1 = the epilogue also emits GroupNorm chunk statistics
};
bool gemm_prepare(const GemmDesc& d, GemmLaunch* l, std::string* err);
void gemm_launch(const GemmLaunch& l, cudaStream_t s);
void gemm_simple_launch(const GemmDesc& d, cudaStream_t s);  // CUDA-core debug path (same semantics)
bool gemm_setup_attributes(std::string* err);
int gemm_plain_bn(int M, int N);                              // tile width of a plain GEMM (see gemm_host.cu)
int gemm_stats_parts(int N, int M);                           // column parts a producer GEMM [M, N] emits
int gemm_set_pair(int on);                                    // CTA-pair kernel on/off; returns the previous value
int gemm_set_sk_min(int k_blocks);                           // stream-K threshold (0 = off); returns the previous value                // opt-in dynamic smem; call once per process/device
// Tile width of a GEGLU projection with N accumulator columns (h | gate halves interleaved per tile by the packing
// kernels): 160 where it divides N (the UNet's 8 C = 2560 / 5120 / 10240: measured 7-11 % faster than 128 - fewer, wider
// MMAs per shared-memory byte and 20 % fewer re-reads of A), else 128 (the stage-1 prior's 16384).
inline int geglu_bn(int N) { return N % 160 == 0 ? 160 : 128; }

// ---- fused GEGLU feed-forward (ffn_fused.cuh; C = 320 only) -----------------------------------------------------
constexpr int FFN_FUSED_C = 320;       // channel width the fused kernel is built for
constexpr int FFN_FUSED_GEGLU_BN = 64; // GEGLU packing width of its W1f / c1 (32 h rows | 32 gate rows per chunk)
struct FfnDesc {
  int dt;                   // DT_F16 | DT_BF16
  int M;                    // rows
  const void* y;            // [M, 320] input = residual
  const float2* stats_in;   // [stats_parts][M] row statistics of y
  int stats_parts;
  float ln_eps;
  const void* w1f;          // [2560, 320] folded (centred, gamma-scaled) weights, GEGLU-packed with width 64
  const float* c1;          // [2560] folded constant vector, same packing
  const void* w2;           // [320, 1280]
  const float* bias2;       // [320]
  void* out;                // [M, 320] (may alias y)
  int pair;                 // 1: CTA-pair kernel
};
struct FfnLaunch {
  FfnMaps maps;
  FfnParams p;
  dim3 grid;
  int dt;
  int pair;  // CTA-pair kernel (library option "ffn_fused" >= 2)
};
bool ffn_prepare(const FfnDesc& d, FfnLaunch* l, std::string* err);
void ffn_launch(const FfnLaunch& l, cudaStream_t s);
bool ffn_setup_attributes(std::string* err);

// ---- attention ---------------------------------------------------------------------------------------
struct AttnDesc {
  int dt;
  const void* q;  // q[(img*S_q + i)*ldq + h*d + c]
  int ldq;
  const void* k;
  const void* v;  // k/v[(img*S_kv + j)*ldkv + h*d + c]
  int ldkv;
  int S_q, S_kv, heads, d, batch;
  void* out;
  int ldo;
};
struct AttnLaunch {
  AttnMaps maps;
  AttnParams p;
  dim3 grid;
  int dpad, dt;
  // short key range (cross_attn_mma_kernel, prior_kernels.cuh): plain pointers instead of tensor maps
  int short_kv, tiles_per_cta;
  AttnDesc desc;
};
bool attn_prepare(const AttnDesc& d, AttnLaunch* l, std::string* err);
void attn_launch(const AttnLaunch& l, cudaStream_t s);
void attn_simple_launch(const AttnDesc& d, cudaStream_t s);
bool attn_setup_attributes(std::string* err);
void temporal_attn_launch(int dt, const void* qkv, void* out, int batch, int frames, int hw, int heads, int d,
                          cudaStream_t s);

}  // namespace rcdm
