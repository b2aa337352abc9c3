// Native model description, weight arena and execution plan of the rich-contextual UNet.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rcdm.h"

namespace rcdm {

struct Mat {  // 16-bit matrix in the weight arena
  size_t off = 0;
  int rows = 0, cols = 0;
};
struct Vec {  // fp32 vector in the weight arena
  size_t off = 0;
  int n = 0;
};

enum SlotKind { SLOT_MAT = 0, SLOT_CONV3 = 1, SLOT_VEC = 2, SLOT_IGNORE = 3 };

struct Slot {  // one state-dict entry of the reference and where / how it lands in the arena
  std::string name;
  int64_t dims[4] = {0, 0, 0, 0};
  int ndim = 0;
  int kind = SLOT_IGNORE;
  size_t dst = 0;
  int ldd = 0, col_off = 0, row_off = 0, geglu_bn = 0, cin = 0;
  bool loaded = false;
};

// LayerNorm folded into the projection that consumes it (gemm_tcgen05.cuh): weights pre-multiplied by gamma plus the
// two epilogue vectors; rebuilt from the raw weights by finalize_weights whenever a state-dict entry changes.
struct LnFold {
  Mat wf;  // [N, K] = W * gamma with every row centred (same row packing as the raw matrix)
  Vec c;   // [frames][N] sum_k (beta[k] + pe[f][k]) W[n, k] + bias[n]
  int frames = 1;
};
struct AttW {
  Mat qkv;  // self / temporal: [3C, C]
  Mat q;    // cross: [C, C]
  Mat kv;   // cross: [2C, ctx]
  Mat out;
  Vec outb;
  LnFold qkv_ln, q_ln;
};
struct ResW {
  int cin = 0, cout = 0;
  bool shortcut = false;
  Vec n1g, n1b, c1b, tb, n2g, n2b, c2b, scb;
  Vec c2beff;  // conv2.bias (+ conv_shortcut.bias), finalised after loading
  Mat c1;      // [cout, 9*cin]
  Mat c2;      // [cout, 9*cout (+ cin when shortcut)]
  int temb_row = 0;  // row offset into the concatenated time_emb_proj matrix / bias vectors
};
struct TfW {
  int C = 0;
  Vec ng, nb, pib, ln1g, ln1b, ln2g, ln2b, ln3g, ln3b, ff1b, ff2b, pob;
  Mat pi, ff1, ff2, po;
  AttW a1, a2;
  LnFold ff1_ln;
  Mat pof;   // [C, 5C] = [po | po * ff2]: proj_out folded over ff.net.2 (no nonlinearity between them), see fold_proj_kernel
  Vec pofb;  // [C] = po * ff2b + pob
};
struct MoW {
  int C = 0;
  Vec ng, nb, pib, lng[4], lnb[4], pe[4], ffng, ffnb, ff1b, ff2b, pob;
  Mat pi, ff1, ff2, po;
  AttW att[4];
  LnFold ff1_ln;
  Mat pof;   // as TfW
  Vec pofb;
};
struct LayerW {
  ResW res;
  bool has_tf = false, has_mo = false;
  TfW tf;
  MoW mo;
};
struct BlockW {
  std::vector<LayerW> layers;
  bool sampler = false;
  Mat sw;
  Vec sb;
  Mat swf;  // up blocks: the sampler weights folded per output parity class, [4 * C, 4 * C] (SEG_UP2)
  int C = 0;
};

using Op = std::function<void(cudaStream_t)>;

struct OpMeta {  // one entry per recorded step op (for rcdm_unet_profile)
  char kind[16];
  double flops;  // algorithmic FLOPs (2*MAC, no padding)
  double bytes;  // algorithmic HBM bytes (inputs + outputs once)
  int m, n, k;   // GEMM-like ops: problem shape (0 otherwise)
};

struct TapInfo {
  size_t off;
  int rows, C;
};

struct rcdm_unet_impl {
  rcdm_unet_config cfg;
  int dt = RCDM_DT_F16;
  // ---- weights
  std::vector<Slot> slots;
  std::unordered_map<std::string, int> slot_index;
  size_t arena_bytes = 0;
  unsigned char* arena = nullptr;
  bool dirty = true;  // derived vectors need (re)finalising
  unsigned char* pack_jobs = nullptr;  // device table of rcdm_unet_load_weights
  size_t pack_jobs_bytes = 0;
  Mat conv_in_w, l1w, l2w, temb_all, conv_out_w;
  Vec conv_in_b, l1b, l2b, temb_static_b, bias_eff_all, c1b_all, tb_all, cno_g, cno_b, conv_out_b;
  int conv_in_kpad = 0, temb_dim = 0, temb_rows = 0;
  std::vector<BlockW> down, up;
  ResW mid_r0, mid_r1;
  TfW mid_tf;
  bool mid_has_mo = false;
  MoW mid_mo;
  // ---- plan
  bool planned = false;
  int B = 0, F = 0, H = 0, W = 0, L = 0;
  unsigned char* ws = nullptr;
  size_t ws_bytes = 0;
  std::vector<Op> ctx_ops, step_ops;
  std::vector<OpMeta> step_meta;
  SkWorkspace sk;  // stream-K partial tiles of this handle's GEMMs (never shared with another handle / stream)
  bool taps_enabled = false;
  std::map<std::string, TapInfo> taps;
  int simple = 0;
  int autotune = 0;  // rcdm_unet_set_option("autotune", 1): plan-time choice of the GEMM tile width / CTA pairing per distinct problem
  std::map<std::string, std::pair<int, int>> tune_cache;  // problem signature -> (tile width, pair)
  int gn_stats = 1;  // GroupNorm statistics from the producing GEMMs' epilogues (0: gn_fused_kernel everywhere)
  size_t gn_acc_bytes = 0;  // size of the accumulator region (from the dry planning pass)
  int ffn_pack64 = 0;  // C = 320 feed-forward weights GEGLU-packed with width 64 for the fused kernel (library option "ffn_fused" at creation)
  int po_fold = 1;  // ff.net.2 + residual -> proj_out + residual as ONE two-segment GEMM on [po | po * ff2] (rcdm_unet_set_option("po_fold", 0): two GEMMs)
  int ln_fold = 1;  // LayerNorm folded into the consuming GEMM (rcdm_unet_set_option("ln_fold", 0): separate layernorm kernels)
  // per-call inputs (read by the recorded ops)
  const void* cur_sample = nullptr;
  int cur_sample_dt = 0;
  const int64_t* cur_t_dev = nullptr;
  float cur_t_host = 0.f;
  const void* cur_ctx = nullptr;
  int cur_ctx_dt = 0;
  void* cur_out = nullptr;
  int cur_out_dt = 0;
  // ---- denoise-loop state
  cudaStream_t loop_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  unsigned char* loop_buf = nullptr;
  size_t loop_buf_bytes = 0;
  cudaGraphExec_t graph_exec = nullptr;
  std::string graph_key;
};

}  // namespace rcdm

struct rcdm_unet : rcdm::rcdm_unet_impl {};
