// UNet3DConditionModel on sm_100a: model construction (state-dict surface), weight repacking, execution plan
// (activation arena + every TMA descriptor resolved once per problem size) and the forward pass.
// Reference: src/models/unet.py:37-462, unet_blocks.py, resnet.py, attention.py, motion_module.py.
//
// Layout: every activation is a channels-last token matrix [(b f y x), C] in the compute dtype, so conv,
// Linear and attention share one layout and none of the reference's einops rearranges exists here.
#include <atomic>
#include <cmath>
#include <cstring>

#include "elementwise.cuh"
#include "internal.h"

namespace rcdm {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int set_err(const std::string& m) {
  g_err = m;
  return 1;
}

#define CUDA_OK(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) return set_err(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

static inline int grid_for(size_t total, int block, int cap = 148 * 16) {
  size_t g = (total + block - 1) / block;
  return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}

// ==========================================================================================
// model construction
// ==========================================================================================
struct Builder {
  rcdm_unet_impl* h;
  size_t top = 0;
  size_t take(size_t bytes) {
    size_t o = top;
    top += (bytes + 255) & ~size_t(255);
    return o;
  }
  Mat mat(int rows, int cols) {
    Mat m;
    m.rows = rows;
    m.cols = cols;
    m.off = take((size_t)rows * cols * 2);
    return m;
  }
  Vec vec(int n) {
    Vec v;
    v.n = n;
    v.off = take((size_t)n * 4);
    return v;
  }
  void slot(const std::string& name, std::initializer_list<int64_t> dims, int kind, size_t dst, int ldd = 0,
            int col_off = 0, int row_off = 0, int geglu_bn = 0, int cin = 0) {
    Slot s;
    s.name = name;
    s.ndim = (int)dims.size();
    int i = 0;
    for (auto d : dims) s.dims[i++] = d;
    s.kind = kind;
    s.dst = dst;
    s.ldd = ldd;
    s.col_off = col_off;
    s.row_off = row_off;
    s.geglu_bn = geglu_bn;
    s.cin = cin;
    h->slot_index[name] = (int)h->slots.size();
    h->slots.push_back(s);
  }
  Vec vslot(const std::string& name, int n) {
    Vec v = vec(n);
    slot(name, {n}, SLOT_VEC, v.off);
    return v;
  }
  Mat lin(const std::string& name, int n, int k) {  // nn.Linear weight [n, k]
    Mat m = mat(n, k);
    slot(name, {n, k}, SLOT_MAT, m.off, k);
    return m;
  }

  LnFold fold(const Mat& w, int frames = 1) {
    LnFold f;
    f.wf = mat(w.rows, w.cols);
    f.c = vec(frames * w.rows);
    f.frames = frames;
    return f;
  }

  void resnet(const std::string& p, int cin, int cout, ResW& r) {
    const int temb = h->temb_dim;
    r.cin = cin;
    r.cout = cout;
    r.shortcut = cin != cout;
    r.n1g = vslot(p + ".norm1.weight", cin);
    r.n1b = vslot(p + ".norm1.bias", cin);
    r.c1 = mat(cout, 9 * cin);
    slot(p + ".conv1.weight", {cout, cin, 3, 3}, SLOT_CONV3, r.c1.off, 9 * cin, 0, 0, 0, cin);
    r.temb_row = h->temb_rows;
    h->temb_rows += cout;
    // conv1.bias / time_emb_proj.{weight,bias} land in the concatenated per-step tables (filled in finish())
    slot(p + ".conv1.bias", {cout}, SLOT_VEC, 0, 0, 0, r.temb_row);
    slot(p + ".time_emb_proj.weight", {cout, temb}, SLOT_MAT, 0, temb, 0, r.temb_row);
    slot(p + ".time_emb_proj.bias", {cout}, SLOT_VEC, 0, 0, 0, r.temb_row);
    r.n2g = vslot(p + ".norm2.weight", cout);
    r.n2b = vslot(p + ".norm2.bias", cout);
    const int k2 = 9 * cout + (r.shortcut ? cin : 0);
    r.c2 = mat(cout, k2);
    slot(p + ".conv2.weight", {cout, cout, 3, 3}, SLOT_CONV3, r.c2.off, k2, 0, 0, 0, cout);
    r.c2b = vslot(p + ".conv2.bias", cout);
    r.c2beff = vec(cout);
    if (r.shortcut) {
      slot(p + ".conv_shortcut.weight", {cout, cin, 1, 1}, SLOT_MAT, r.c2.off, k2, 9 * cout);
      r.scb = vslot(p + ".conv_shortcut.bias", cout);
    }
  }
  void attn(const std::string& p, int C, int kvdim, bool cross, AttW& a) {
    if (!cross) {
      a.qkv = mat(3 * C, C);
      slot(p + ".to_q.weight", {C, C}, SLOT_MAT, a.qkv.off, C, 0, 0);
      slot(p + ".to_k.weight", {C, C}, SLOT_MAT, a.qkv.off, C, 0, C);
      slot(p + ".to_v.weight", {C, C}, SLOT_MAT, a.qkv.off, C, 0, 2 * C);
    } else {
      a.q = lin(p + ".to_q.weight", C, C);
      a.kv = mat(2 * C, kvdim);
      slot(p + ".to_k.weight", {C, kvdim}, SLOT_MAT, a.kv.off, kvdim, 0, 0);
      slot(p + ".to_v.weight", {C, kvdim}, SLOT_MAT, a.kv.off, kvdim, 0, C);
    }
    a.out = lin(p + ".to_out.0.weight", C, C);
    a.outb = vslot(p + ".to_out.0.bias", C);
  }
  void ff(const std::string& p, int C, Mat& ff1, Vec& ff1b, Mat& ff2, Vec& ff2b) {
    // GEGLU row interleave: the tile width of the GEGLU GEMM - or, for the 320-channel blocks, the chunk width of the
    // fused feed-forward kernel (ffn_fused.cuh)
    const int gbn = (C == FFN_FUSED_C && h->ffn_pack64) ? FFN_FUSED_GEGLU_BN : geglu_bn(8 * C);
    ff1 = mat(8 * C, C);
    slot(p + ".net.0.proj.weight", {8 * C, C}, SLOT_MAT, ff1.off, C, 0, 0, gbn);
    ff1b = vec(8 * C);
    slot(p + ".net.0.proj.bias", {8 * C}, SLOT_VEC, ff1b.off, 0, 0, 0, gbn);
    ff2 = lin(p + ".net.2.weight", C, 4 * C);
    ff2b = vslot(p + ".net.2.bias", C);
  }
  void transformer(const std::string& p, int C, TfW& t) {
    const std::string b = p + ".transformer_blocks.0";
    t.C = C;
    t.ng = vslot(p + ".norm.weight", C);
    t.nb = vslot(p + ".norm.bias", C);
    t.pi = mat(C, C);
    slot(p + ".proj_in.weight", {C, C, 1, 1}, SLOT_MAT, t.pi.off, C);
    t.pib = vslot(p + ".proj_in.bias", C);
    attn(b + ".attn1", C, C, false, t.a1);
    t.a1.qkv_ln = fold(t.a1.qkv);
    t.ln1g = vslot(b + ".norm1.weight", C);
    t.ln1b = vslot(b + ".norm1.bias", C);
    attn(b + ".attn2", C, h->cfg.cross_attention_dim, true, t.a2);
    t.a2.q_ln = fold(t.a2.q);
    t.ln2g = vslot(b + ".norm2.weight", C);
    t.ln2b = vslot(b + ".norm2.bias", C);
    ff(b + ".ff", C, t.ff1, t.ff1b, t.ff2, t.ff2b);
    t.ff1_ln = fold(t.ff1);
    t.ln3g = vslot(b + ".norm3.weight", C);
    t.ln3b = vslot(b + ".norm3.bias", C);
    t.po = mat(C, C);
    slot(p + ".proj_out.weight", {C, C, 1, 1}, SLOT_MAT, t.po.off, C);
    t.pob = vslot(p + ".proj_out.bias", C);
    t.pof = mat(C, 5 * C);
    t.pofb = vec(C);
  }
  void motion(const std::string& p0, int C, MoW& m) {
    const std::string p = p0 + ".temporal_transformer";
    const std::string b = p + ".transformer_blocks.0";
    const int na = h->cfg.motion_attn_blocks, ml = h->cfg.motion_max_len;
    m.C = C;
    m.ng = vslot(p + ".norm.weight", C);
    m.nb = vslot(p + ".norm.bias", C);
    slot(p + ".prior_norm.weight", {C}, SLOT_IGNORE, 0);  // stage-1 only (motion_module.py:150-153)
    slot(p + ".prior_norm.bias", {C}, SLOT_IGNORE, 0);
    m.pi = lin(p + ".proj_in.weight", C, C);
    m.pib = vslot(p + ".proj_in.bias", C);
    for (int i = 0; i < na; ++i) {
      const std::string a = b + ".attention_blocks." + std::to_string(i);
      attn(a, C, C, false, m.att[i]);
      m.att[i].qkv_ln = fold(m.att[i].qkv, ml);
      m.pe[i] = vec(ml * C);
      slot(a + ".pos_encoder.pe", {1, ml, C}, SLOT_VEC, m.pe[i].off);
    }
    for (int i = 0; i < na; ++i) {
      m.lng[i] = vslot(b + ".norms." + std::to_string(i) + ".weight", C);
      m.lnb[i] = vslot(b + ".norms." + std::to_string(i) + ".bias", C);
    }
    ff(b + ".ff", C, m.ff1, m.ff1b, m.ff2, m.ff2b);
    m.ff1_ln = fold(m.ff1);
    m.ffng = vslot(b + ".ff_norm.weight", C);
    m.ffnb = vslot(b + ".ff_norm.bias", C);
    m.po = lin(p + ".proj_out.weight", C, C);
    m.pob = vslot(p + ".proj_out.bias", C);
    m.pof = mat(C, 5 * C);
    m.pofb = vec(C);
  }
};

static int build_model(rcdm_unet_impl* h) {
  const rcdm_unet_config& c = h->cfg;
  if (c.num_blocks < 1 || c.num_blocks > RCDM_MAX_BLOCKS) return set_err("num_blocks out of range");
  if (c.compute_dtype != RCDM_DT_F16 && c.compute_dtype != RCDM_DT_BF16)
    return set_err("compute_dtype must be RCDM_DT_F16 or RCDM_DT_BF16 (no fp32 tensor-core path)");
  if (c.motion_attn_blocks > 4) return set_err("at most 4 temporal attention blocks");
  if (c.motion_max_len > 5) return set_err("temporal_position_encoding_max_len > 5 is not supported");
  for (int i = 0; i < c.num_blocks; ++i) {
    const int C = c.block_out_channels[i];
    if (C % 64 != 0 || C % c.norm_num_groups != 0 || C % (8 * c.attention_heads) != 0)
      return set_err("block_out_channels must be multiples of 64, of norm_num_groups and of 8*heads");
  }
  if (c.cross_attention_dim % 8 != 0) return set_err("cross_attention_dim must be a multiple of 8");
  h->dt = c.compute_dtype;
  Builder b{h};
  const int c0 = c.block_out_channels[0];
  h->temb_dim = 4 * c0;
  h->temb_rows = 0;
  h->conv_in_kpad = (9 * c.in_channels + 7) / 8 * 8;
  h->conv_in_w = b.mat(c0, h->conv_in_kpad);
  b.slot("conv_in.weight", {c0, c.in_channels, 3, 3}, SLOT_CONV3, h->conv_in_w.off, h->conv_in_kpad, 0, 0, 0,
         c.in_channels);
  h->conv_in_b = b.vslot("conv_in.bias", c0);
  h->l1w = b.lin("time_embedding.linear_1.weight", h->temb_dim, c0);
  h->l1b = b.vslot("time_embedding.linear_1.bias", h->temb_dim);
  h->l2w = b.lin("time_embedding.linear_2.weight", h->temb_dim, h->temb_dim);
  h->l2b = b.vslot("time_embedding.linear_2.bias", h->temb_dim);

  h->down.resize(c.num_blocks);
  int out_c = c0;
  for (int i = 0; i < c.num_blocks; ++i) {
    const int in_c = out_c;
    out_c = c.block_out_channels[i];
    BlockW& blk = h->down[i];
    blk.C = out_c;
    blk.layers.resize(c.layers_per_block);
    const std::string p = "down_blocks." + std::to_string(i);
    if (c.down_has_attn[i])
      for (int j = 0; j < c.layers_per_block; ++j) {
        blk.layers[j].has_tf = true;
        b.transformer(p + ".attentions." + std::to_string(j), out_c, blk.layers[j].tf);
      }
    for (int j = 0; j < c.layers_per_block; ++j)
      b.resnet(p + ".resnets." + std::to_string(j), j == 0 ? in_c : out_c, out_c, blk.layers[j].res);
    if (c.use_motion_module && c.motion_down[i])
      for (int j = 0; j < c.layers_per_block; ++j) {
        blk.layers[j].has_mo = true;
        b.motion(p + ".motion_modules." + std::to_string(j), out_c, blk.layers[j].mo);
      }
    blk.sampler = i != c.num_blocks - 1;
    if (blk.sampler) {
      blk.sw = b.mat(out_c, 9 * out_c);
      b.slot(p + ".downsamplers.0.conv.weight", {out_c, out_c, 3, 3}, SLOT_CONV3, blk.sw.off, 9 * out_c, 0, 0, 0, out_c);
      blk.sb = b.vslot(p + ".downsamplers.0.conv.bias", out_c);
    }
  }
  h->up.resize(c.num_blocks);
  out_c = c.block_out_channels[c.num_blocks - 1];
  for (int i = 0; i < c.num_blocks; ++i) {
    const int prev = out_c;
    out_c = c.block_out_channels[c.num_blocks - 1 - i];
    const int in_idx = c.num_blocks - 1 - (i + 1 < c.num_blocks ? i + 1 : c.num_blocks - 1);
    const int in_c = c.block_out_channels[in_idx];
    BlockW& blk = h->up[i];
    blk.C = out_c;
    const int nl = c.layers_per_block + 1;
    blk.layers.resize(nl);
    const std::string p = "up_blocks." + std::to_string(i);
    if (c.up_has_attn[i])
      for (int j = 0; j < nl; ++j) {
        blk.layers[j].has_tf = true;
        b.transformer(p + ".attentions." + std::to_string(j), out_c, blk.layers[j].tf);
      }
    for (int j = 0; j < nl; ++j) {
      const int skip = (j == nl - 1) ? in_c : out_c;
      const int rin = (j == 0) ? prev : out_c;
      b.resnet(p + ".resnets." + std::to_string(j), rin + skip, out_c, blk.layers[j].res);
    }
    if (c.use_motion_module && c.motion_up[i])
      for (int j = 0; j < nl; ++j) {
        blk.layers[j].has_mo = true;
        b.motion(p + ".motion_modules." + std::to_string(j), out_c, blk.layers[j].mo);
      }
    blk.sampler = i != c.num_blocks - 1;
    if (blk.sampler) {
      blk.sw = b.mat(out_c, 9 * out_c);
      b.slot(p + ".upsamplers.0.conv.weight", {out_c, out_c, 3, 3}, SLOT_CONV3, blk.sw.off, 9 * out_c, 0, 0, 0, out_c);
      blk.sb = b.vslot(p + ".upsamplers.0.conv.bias", out_c);
      blk.swf = b.mat(4 * out_c, 4 * out_c);
    }
  }
  const int mc = c.block_out_channels[c.num_blocks - 1];
  b.transformer("mid_block.attentions.0", mc, h->mid_tf);
  b.resnet("mid_block.resnets.0", mc, mc, h->mid_r0);
  b.resnet("mid_block.resnets.1", mc, mc, h->mid_r1);
  h->mid_has_mo = c.use_motion_module && c.motion_mid;
  if (h->mid_has_mo) b.motion("mid_block.motion_modules.0", mc, h->mid_mo);
  h->cno_g = b.vslot("conv_norm_out.weight", c0);
  h->cno_b = b.vslot("conv_norm_out.bias", c0);
  h->conv_out_w = b.mat(c.out_channels, 9 * c0);
  b.slot("conv_out.weight", {c.out_channels, c0, 3, 3}, SLOT_CONV3, h->conv_out_w.off, 9 * c0, 0, 0, 0, c0);
  h->conv_out_b = b.vslot("conv_out.bias", c.out_channels);

  // concatenated per-step time-embedding tables (all resnets at once)
  h->temb_all = b.mat(h->temb_rows, h->temb_dim);
  h->c1b_all = b.vec(h->temb_rows);
  h->tb_all = b.vec(h->temb_rows);
  h->temb_static_b = b.vec(h->temb_rows);
  h->bias_eff_all = b.vec(h->temb_rows);
  for (auto& s : h->slots) {
    const std::string& n = s.name;
    auto ends = [&](const char* suf) {
      const size_t l = strlen(suf);
      return n.size() >= l && n.compare(n.size() - l, l, suf) == 0;
    };
    if (ends(".time_emb_proj.weight")) s.dst = h->temb_all.off;
    else if (ends(".time_emb_proj.bias")) s.dst = h->tb_all.off + (size_t)s.row_off * 4, s.row_off = 0;
    else if (ends(".conv1.bias")) s.dst = h->c1b_all.off + (size_t)s.row_off * 4, s.row_off = 0;
  }
  h->arena_bytes = b.top;
  return 0;
}

// ==========================================================================================
// weights
// ==========================================================================================
template <typename T>
static void pack_dispatch(rcdm_unet_impl* h, const Slot& s, const void* src, int src_dt, cudaStream_t st) {
  if (s.kind == SLOT_VEC) {
    int n = 1;
    for (int i = 0; i < s.ndim; ++i) n *= (int)s.dims[i];
    pack_vec_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, src_dt, reinterpret_cast<float*>(h->arena + s.dst), n,
                                                     s.row_off, s.geglu_bn, 0);
  } else {
    const int N = (int)s.dims[0];
    int K = 1;
    for (int i = 1; i < s.ndim; ++i) K *= (int)s.dims[i];
    const bool conv3 = s.kind == SLOT_CONV3;
    pack_weight_kernel<T><<<grid_for((size_t)N * K, 256), 256, 0, st>>>(
        src, src_dt, reinterpret_cast<T*>(h->arena + s.dst), N, K, s.ldd, s.col_off, s.row_off, conv3 ? 1 : 0, s.cin,
        s.geglu_bn);
  }
  g_launches++;
}

__global__ void add_vec_kernel(const float* a, const float* b, float* o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + (b ? b[i] : 0.f);
}

static void finalize_res(rcdm_unet_impl* h, const ResW& r, cudaStream_t st) {
  auto f = [&](const Vec& v) { return reinterpret_cast<float*>(h->arena + v.off); };
  add_vec_kernel<<<(r.cout + 255) / 256, 256, 0, st>>>(f(r.c2b), r.shortcut ? f(r.scb) : nullptr, f(r.c2beff), r.cout);
}

static void fold_one(rcdm_unet_impl* h, const Mat& w, const LnFold& lf, const Vec& g, const Vec& b, const Vec* pe,
                     const Vec* bias, cudaStream_t st) {
  auto f = [&](const Vec& v) { return reinterpret_cast<float*>(h->arena + v.off); };
  const int N = w.rows, K = w.cols;
  const int blocks = (N * 32 + 255) / 256;
  if (h->dt == DT_F16)
    fold_ln_kernel<__half><<<blocks, 256, 0, st>>>(reinterpret_cast<const __half*>(h->arena + w.off),
                                                   reinterpret_cast<__half*>(h->arena + lf.wf.off), f(g), f(b),
                                                   pe ? f(*pe) : nullptr, bias ? f(*bias) : nullptr, f(lf.c), N, K,
                                                   lf.frames);
  else
    fold_ln_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(h->arena + w.off), reinterpret_cast<__nv_bfloat16*>(h->arena + lf.wf.off),
        f(g), f(b), pe ? f(*pe) : nullptr, bias ? f(*bias) : nullptr, f(lf.c), N, K, lf.frames);
}
// proj_out folded over the feed-forward's second Linear.  Transformer3DModel / TemporalTransformer3DModel end with
//   y2 = y + ff2(g) + b2;  x = x + po(y2) + bp        (attention.py:347-359,514; motion_module.py:170-180,243)
// with nothing non-linear in between, so x = x + [y | g] [Wp | Wp W2]^T + (Wp b2 + bp): one GEMM over two K segments, the
// intermediate y2 (one write + one read of the hidden state, one launch) never exists.
//   wf [C, 5C]: columns 0..C-1 = Wp, columns C.. = Wp W2 (fp32 accumulation over the 16-bit weights, rounded once)
//   cf [C]    : Wp b2 + bp (fp32)
constexpr int FOLDP_NB = 8;  // output rows per block (each W2 element is read once per FOLDP_NB rows)
template <typename T>
__global__ void fold_proj_kernel(const T* __restrict__ Wp, const T* __restrict__ W2, const float* __restrict__ b2,
                                 const float* __restrict__ bp, T* __restrict__ wf, float* __restrict__ cf, int C) {
  const int J = 4 * C, n0 = blockIdx.y * FOLDP_NB;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float wp_s[FOLDP_NB][64];
  float acc[FOLDP_NB];
#pragma unroll
  for (int i = 0; i < FOLDP_NB; ++i) acc[i] = 0.f;
  for (int c0 = 0; c0 < C; c0 += 64) {
    __syncthreads();
    for (int e = threadIdx.x; e < FOLDP_NB * 64; e += blockDim.x) {
      const int i = e / 64, c = c0 + e % 64;
      wp_s[i][e % 64] = (n0 + i < C && c < C) ? DT<T>::to_f(Wp[(size_t)(n0 + i) * C + c]) : 0.f;
    }
    __syncthreads();
    if (j < J) {
      const int cm = min(64, C - c0);
      for (int c = 0; c < cm; ++c) {
        const float w2 = DT<T>::to_f(W2[(size_t)(c0 + c) * J + j]);
#pragma unroll
        for (int i = 0; i < FOLDP_NB; ++i) acc[i] = fmaf(wp_s[i][c], w2, acc[i]);
      }
    }
  }
  if (j < J) {
#pragma unroll
    for (int i = 0; i < FOLDP_NB; ++i)
      if (n0 + i < C) wf[(size_t)(n0 + i) * 5 * C + C + j] = DT<T>::from_f(acc[i]);
  }
  // the unchanged proj_out columns and the constant vector: by the blocks of the first column slab
  if (blockIdx.x == 0) {
    for (int e = threadIdx.x; e < FOLDP_NB * C; e += blockDim.x) {
      const int i = e / C, c = e % C;
      if (n0 + i < C) wf[(size_t)(n0 + i) * 5 * C + c] = Wp[(size_t)(n0 + i) * C + c];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < FOLDP_NB; i += blockDim.x >> 5) {
      if (n0 + i >= C) continue;
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s = fmaf(DT<T>::to_f(Wp[(size_t)(n0 + i) * C + c]), b2[c], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) cf[n0 + i] = s + bp[n0 + i];
    }
  }
}
void fold_proj_launch(int dt, const void* wp, const void* w2, const float* b2, const float* bp, void* wf, float* cf, int C,
                      cudaStream_t st) {
  const dim3 grid((4 * C + 255) / 256, (C + FOLDP_NB - 1) / FOLDP_NB);
  if (dt == DT_F16)
    fold_proj_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(wp), reinterpret_cast<const __half*>(w2),
                                                   b2, bp, reinterpret_cast<__half*>(wf), cf, C);
  else
    fold_proj_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(wp),
                                                          reinterpret_cast<const __nv_bfloat16*>(w2), b2, bp,
                                                          reinterpret_cast<__nv_bfloat16*>(wf), cf, C);
}
static void fold_proj(rcdm_unet_impl* h, const Mat& po, const Mat& ff2, const Vec& ff2b, const Vec& pob, const Mat& pof,
                      const Vec& pofb, cudaStream_t st) {
  auto f = [&](const Vec& v) { return reinterpret_cast<float*>(h->arena + v.off); };
  fold_proj_launch(h->dt, h->arena + po.off, h->arena + ff2.off, f(ff2b), f(pob), h->arena + pof.off, f(pofb), po.rows, st);
}
static void fold_tf(rcdm_unet_impl* h, const TfW& t, cudaStream_t st) {
  fold_proj(h, t.po, t.ff2, t.ff2b, t.pob, t.pof, t.pofb, st);
  fold_one(h, t.a1.qkv, t.a1.qkv_ln, t.ln1g, t.ln1b, nullptr, nullptr, st);
  fold_one(h, t.a2.q, t.a2.q_ln, t.ln2g, t.ln2b, nullptr, nullptr, st);
  fold_one(h, t.ff1, t.ff1_ln, t.ln3g, t.ln3b, nullptr, &t.ff1b, st);
}
static void fold_mo(rcdm_unet_impl* h, const MoW& m, cudaStream_t st) {
  fold_proj(h, m.po, m.ff2, m.ff2b, m.pob, m.pof, m.pofb, st);
  for (int i = 0; i < h->cfg.motion_attn_blocks; ++i)
    fold_one(h, m.att[i].qkv, m.att[i].qkv_ln, m.lng[i], m.lnb[i], &m.pe[i], nullptr, st);
  fold_one(h, m.ff1, m.ff1_ln, m.ffng, m.ffnb, nullptr, &m.ff1b, st);
}

static void finalize_weights(rcdm_unet_impl* h, cudaStream_t st) {
  auto f = [&](const Vec& v) { return reinterpret_cast<float*>(h->arena + v.off); };
  for (auto* blks : {&h->down, &h->up})
    for (auto& blk : *blks)
      for (auto& l : blk.layers) {
        if (l.has_tf) fold_tf(h, l.tf, st);
        if (l.has_mo) fold_mo(h, l.mo, st);
      }
  fold_tf(h, h->mid_tf, st);
  if (h->mid_has_mo) fold_mo(h, h->mid_mo, st);
  for (auto& blk : h->up)
    if (blk.sampler) {  // Upsample3D folded into its conv: per-parity-class 2x2 weights
      const size_t total = (size_t)16 * blk.C * blk.C;
      if (h->dt == DT_F16)
        fold_upsample_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>(
            reinterpret_cast<const __half*>(h->arena + blk.sw.off), reinterpret_cast<__half*>(h->arena + blk.swf.off), blk.C, blk.C);
      else
        fold_upsample_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(h->arena + blk.sw.off),
            reinterpret_cast<__nv_bfloat16*>(h->arena + blk.swf.off), blk.C, blk.C);
    }
  for (auto& blk : h->down)
    for (auto& l : blk.layers) finalize_res(h, l.res, st);
  for (auto& blk : h->up)
    for (auto& l : blk.layers) finalize_res(h, l.res, st);
  finalize_res(h, h->mid_r0, st);
  finalize_res(h, h->mid_r1, st);
  add_vec_kernel<<<(h->temb_rows + 255) / 256, 256, 0, st>>>(f(h->c1b_all), f(h->tb_all), f(h->temb_static_b),
                                                             h->temb_rows);
  h->dirty = false;
}

// ==========================================================================================
// plan
// ==========================================================================================
constexpr size_t NO_GN = ~size_t(0);
struct Act {
  size_t off = 0;
  int C = 0, H = 0, W = 0;
  size_t gn = NO_GN;  // workspace offset of this tensor's GroupNorm chunk accumulators (written by its producer's epilogue)
};

struct Planner {
  rcdm_unet_impl* h;
  bool dry;
  std::vector<Op>* ops;
  std::map<size_t, size_t> free_list;
  size_t top = 0, peak = 0;
  std::string err;
  bool failed = false;
  size_t gn_scratch = 0;  // persistent scratch for GroupNorm (counters stay zero between launches)
  // GroupNorm statistics from the producers' epilogues: one accumulator block per normalised tensor, all inside one
  // region that the first op of a step zeroes (sized by the dry pass)
  size_t gn_region = 0, gn_used = 0;
  bool gn_stats_on() const { return h->gn_stats && !h->simple; }
  size_t new_gn(int C, int H, int W) {
    if (!gn_stats_on() || !gemm_gn_stats_ok(h->B * h->F * H * W, C, H * W)) return NO_GN;
    const size_t off = gn_region + gn_used;
    gn_used += (gemm_gn_acc_bytes(h->B * h->F, C) + 255) & ~size_t(255);
    return off;
  }
  unsigned long long* gnp(size_t off) const { return reinterpret_cast<unsigned long long*>(h->ws + off); }

  size_t alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    for (auto it = free_list.begin(); it != free_list.end(); ++it) {
      if (it->second >= bytes) {
        const size_t off = it->first, sz = it->second;
        free_list.erase(it);
        if (sz > bytes) free_list[off + bytes] = sz - bytes;
        return off;
      }
    }
    const size_t off = top;
    top += bytes;
    if (top > peak) peak = top;
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    auto it = free_list.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != free_list.end() && it->first + it->second == nx->first) {
      it->second += nx->second;
      free_list.erase(nx);
    }
    if (it != free_list.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) {
        pv->second += it->second;
        free_list.erase(it);
        it = pv;
      }
    }
    if (it->first + it->second == top) {
      top = it->first;
      free_list.erase(it);
    }
  }
  int rows(const Act& a) const { return h->B * h->F * a.H * a.W; }
  size_t abytes(const Act& a) const { return (size_t)rows(a) * a.C * 2; }
  Act new_act(int C, int H, int W) {
    Act a;
    a.C = C;
    a.H = H;
    a.W = W;
    a.off = alloc(abytes(a));
    return a;
  }
  void free_act(const Act& a) { release(a.off, abytes(a)); }
  void* p(size_t off) const { return h->ws + off; }
  const void* wm(const Mat& m) const { return h->arena + m.off; }
  const float* wv(const Vec& v) const { return reinterpret_cast<const float*>(h->arena + v.off); }
  void fail(const std::string& m) {
    if (!failed) err = m;
    failed = true;
  }
  void push(Op op, const char* kind = "misc", double flops = 0.0, double bytes = 0.0, int pm = 0, int pn = 0, int pk = 0) {
    if (dry) return;
    ops->push_back(std::move(op));
    if (ops == &h->step_ops) {
      OpMeta m;
      memset(&m, 0, sizeof m);
      strncpy(m.kind, kind, sizeof(m.kind) - 1);
      m.flops = flops;
      m.bytes = bytes;
      m.m = pm;
      m.n = pn;
      m.k = pk;
      h->step_meta.push_back(m);
    }
  }

  // ---------------- op emitters ----------------
  void gemm(GemmDesc d) {
    if (dry || failed) return;
    d.dt = h->dt;
    // algorithmic FLOPs: 2*M*N*K with K = sum of segment channels (x9 for 3x3 taps), no padding
    double k_alg = 0;
    bool conv = false;
    for (int i = 0; i < d.nseg; ++i) {
      // algorithmic K: a folded upsample (SEG_UP2) executes 4 taps but stands for the reference's 9
      k_alg += d.seg[i].mode == SEG_PLAIN ? d.seg[i].C : 9.0 * d.seg[i].C;
      conv |= d.seg[i].mode != SEG_PLAIN;
    }
    if (d.seg[0].mode == SEG_PLAIN && d.nseg == 1 && d.seg[0].C == h->conv_in_kpad && d.w == wm(h->conv_in_w))
      k_alg = 9.0 * h->cfg.in_channels;
    const double flops = 2.0 * d.M * d.N * k_alg;
    const double bytes = 2.0 * ((double)d.M * k_alg / (conv ? 9.0 : 1.0) + (double)d.N * k_alg +
                                (double)d.M * (d.geglu ? d.N / 2 : d.N));
    const char* kind = conv ? "conv3x3" : (d.geglu ? "gemm_geglu" : "gemm");
    if (h->simple) {
      push([d](cudaStream_t s) {
        gemm_simple_launch(d, s);
        g_launches++;
      }, kind, flops, bytes, d.M, d.N, (int)k_alg);
      return;
    }
    GemmLaunch l;
    std::string e;
    d.sk = &h->sk;  // the handle's own stream-K workspace: its plan runs on one stream at a time
    if (h->autotune) tune(d);
    if (!gemm_prepare(d, &l, &e)) return fail(e);
    push([l](cudaStream_t s) {
      gemm_launch(l, s);
      g_launches++;
    }, kind, flops, bytes, d.M, d.N, (int)k_alg);
  }
  // Plan-time autotuning: every distinct GEMM problem is timed once (CUDA events, real buffers of the arena) with
  // each admissible tile width x {single CTA, CTA pair}; the winner is cached by problem signature.  The K summation
  // order of an output element does not depend on these two choices, so results stay bitwise identical; the stream-K
  // decision (which does change the order) stays with the deterministic heuristic.
  void tune(GemmDesc& d) {
    char key[160];
    snprintf(key, sizeof key, "%d/%d/%d/%d/%d%d%d/%d/%d/%d/%d/%d/%d", d.M, d.N, d.Ktot, d.nseg, d.seg[0].mode, d.seg[1].mode,
             d.seg[2].mode, d.geglu, d.res != nullptr, d.stats_out != nullptr, d.stats_in != nullptr, d.Ho, d.bias != nullptr);
    auto it = h->tune_cache.find(key);
    if (it == h->tune_cache.end()) {
      std::vector<int> bns;
      if (d.geglu) bns = {geglu_bn(d.N)};
      else if (d.stats_out) bns = {0};  // the consumer reads 2 * N-tiles statistic parts: keep the heuristic width
      else {
        bns = {64, 128};
        if (d.N % 160 == 0) bns.push_back(160);
        if (d.N <= 64) bns = {64};
      }
      cudaStream_t ts;
      cudaEvent_t e0, e1;
      cudaStreamCreateWithFlags(&ts, cudaStreamNonBlocking);
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      float best = 1e30f;
      std::pair<int, int> pick{0, 0};
      for (int bn : bns)
        for (int pr = 0; pr < 2; ++pr) {
          GemmDesc c = d;
          c.force_bn = bn;
          c.no_pair = pr ? 0 : 1;
          c.force_pair = pr;
          GemmLaunch l;
          std::string err;
          if (!gemm_prepare(c, &l, &err)) continue;
          if (pr && !l.pair) continue;  // pairing not applicable (single M tile)
          for (int w = 0; w < 2; ++w) gemm_launch(l, ts);
          cudaEventRecord(e0, ts);
          for (int r = 0; r < 6; ++r) gemm_launch(l, ts);
          cudaEventRecord(e1, ts);
          if (cudaEventSynchronize(e1) != cudaSuccess) continue;
          float ms = 0.f;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) {
            best = ms;
            pick = {bn, pr};
          }
        }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      cudaStreamDestroy(ts);
      it = h->tune_cache.emplace(key, pick).first;
    }
    d.force_bn = it->second.first;
    d.no_pair = it->second.second ? 0 : 1;
    d.force_pair = it->second.second;
  }
  // plain GEMM: out[M,N] = A[M,K] W^T (+bias)(+res)
  // ln != nullptr: the input is consumed through a folded LayerNorm (statistics at stats_off, emitted by the GEMM that
  // produced the input); emit_stats: this GEMM's output feeds a folded LayerNorm -> write its row statistics
  void linear(size_t a_off, int M, int K, const Mat& w, const Vec* bias, size_t out_off, int ldo, const size_t* res_off,
              int geglu = 0, const LnFold* ln = nullptr, size_t stats_off = 0, bool emit_stats = false,
              int rows_per_frame = 1, size_t gn_off = NO_GN, int gn_hw = 0) {
    GemmDesc d;
    memset(&d, 0, sizeof d);
    if (gn_off != NO_GN) {
      d.gn_acc = gnp(gn_off);
      d.gn_hw = gn_hw;
    }
    d.M = M;
    d.N = w.rows;
    d.nseg = 1;
    d.seg[0] = ASeg{SEG_PLAIN, p(a_off), K, K, 0, 0, 0};
    if (ln) {
      d.stats_in = reinterpret_cast<const float2*>(p(stats_off));
      d.stats_parts = gemm_stats_parts(K, M);
      d.ln_c = wv(ln->c);
      d.ln_frames = ln->frames > 1 ? h->F : 1;
      d.ln_rows_per_frame = rows_per_frame;
      d.ln_eps = 1e-5f;
      bias = nullptr;  // folded into c
    }
    if (emit_stats) d.stats_out = reinterpret_cast<float2*>(p(stats_off));
    d.w = ln ? wm(ln->wf) : wm(w);
    d.Ktot = w.cols;
    d.w_rows = w.rows;
    d.out = p(out_off);
    d.ldo = ldo;
    d.bias = bias ? wv(*bias) : nullptr;
    d.res = res_off ? p(*res_off) : nullptr;
    d.ldr = ldo;
    d.geglu = geglu;
    if (geglu && K == FFN_FUSED_C && h->ffn_pack64) d.force_bn = FFN_FUSED_GEGLU_BN;  // (packed for the fused kernel)
    if (K != w.cols) return fail("linear: K mismatch");
    gemm(d);
  }
  // x <- x + [y | g] [po | po ff2]^T + (po ff2b + pob): ff.net.2 (+ residual) and proj_out (+ residual) as one launch
  // over two K segments (fold_proj_kernel); the output's GroupNorm statistics come from this epilogue as before
  bool po_fold_ok(int C) const { return h->po_fold && !h->simple && C % 64 == 0; }
  void proj_out_folded(size_t y_off, size_t g_off, int M, int C, const Mat& pof, const Vec& pofb, size_t x_off,
                       size_t gn_off, int gn_hw) {
    GemmDesc d;
    memset(&d, 0, sizeof d);
    if (gn_off != NO_GN) {
      d.gn_acc = gnp(gn_off);
      d.gn_hw = gn_hw;
    }
    d.M = M;
    d.N = C;
    d.nseg = 2;
    d.seg[0] = ASeg{SEG_PLAIN, p(y_off), C, C, 0, 0, 0};
    d.seg[1] = ASeg{SEG_PLAIN, p(g_off), 4 * C, 4 * C, 0, 0, 0};
    d.w = wm(pof);
    d.Ktot = 5 * C;
    d.w_rows = C;
    d.out = p(x_off);
    d.ldo = C;
    d.bias = wv(pofb);
    d.res = p(x_off);
    d.ldr = C;
    gemm(d);
  }
  // fused GEGLU feed-forward of a 320-channel block (ffn_fused.cuh): y <- y + GEGLU(LN(y) W1^T + b1) W2^T + b2
  bool ffn_fused_ok(int C, bool fold) const { return fold && C == FFN_FUSED_C && h->ffn_pack64 && opt(OPT_FFN_FUSED); }
  void ffn_fused(size_t y_off, int M, const LnFold& ln, size_t stats_off, const Mat& w2, const Vec& b2) {
    if (dry || failed) return;
    FfnDesc d;
    memset(&d, 0, sizeof d);
    d.dt = h->dt;
    d.M = M;
    d.y = p(y_off);
    d.stats_in = reinterpret_cast<const float2*>(p(stats_off));
    d.stats_parts = gemm_stats_parts(FFN_FUSED_C, M);
    d.ln_eps = 1e-5f;
    d.w1f = wm(ln.wf);
    d.c1 = wv(ln.c);
    d.w2 = wm(w2);
    d.bias2 = wv(b2);
    d.out = p(y_off);
    d.pair = opt(OPT_FFN_FUSED) >= 2 ? 1 : 0;
    FfnLaunch l;
    std::string e;
    if (!ffn_prepare(d, &l, &e)) return fail(e);
    const double C = FFN_FUSED_C, J = 4.0 * FFN_FUSED_C;
    push([l](cudaStream_t s) {
      ffn_launch(l, s);
      g_launches++;
    }, "ffn_fused", 2.0 * M * (2.0 * J * C + C * J), 2.0 * (2.0 * M * C + 2.0 * J * C + C * J), M, (int)C, (int)J);
  }
  void groupnorm(const Act& x0, const Act* x1, const Vec& g, const Vec& b, float eps, bool per_frame, bool silu,
                 size_t out_off) {
    if (dry || failed) return;
    const int hw = x0.H * x0.W;
    GnLaunch l;
    // algorithmic bytes: one read + one write of the tensor
    const double bytes = 2.0 * rows(x0) * (x0.C + (x1 ? x1->C : 0)) * 2.0;
    if (x0.gn != NO_GN && (!x1 || x1->gn != NO_GN) && x0.C + (x1 ? x1->C : 0) <= 3072) {
      // statistics were emitted by the producing GEMMs' epilogues: one streaming normalise (+SiLU) pass
      gn_configure_from_stats(&l, h->dt, p(x0.off), x0.C, gnp(x0.gn), x1 ? p(x1->off) : nullptr, x1 ? x1->C : 0,
                              x1 ? gnp(x1->gn) : nullptr, rows(x0), per_frame ? hw : h->F * hw, hw,
                              h->cfg.norm_num_groups, eps, wv(g), wv(b), p(out_off), silu ? 1 : 0);
      push([l](cudaStream_t s) { gn_run(l, s); }, "groupnorm", 0.0, bytes, rows(x0), x0.C + (x1 ? x1->C : 0),
           per_frame ? 1 : h->F);
      return;
    }
    gn_configure(&l, h->dt, p(x0.off), x0.C, x1 ? p(x1->off) : nullptr, x1 ? x1->C : 0, rows(x0),
                 per_frame ? hw : h->F * hw, h->cfg.norm_num_groups, eps, wv(g), wv(b), p(out_off), silu ? 1 : 0,
                 p(gn_scratch));
    push([l](cudaStream_t s) { gn_run(l, s); }, "groupnorm", 0.0, bytes, rows(x0), x0.C + (x1 ? x1->C : 0),
         per_frame ? -1 : -h->F);  // k < 0: statistics computed by the GroupNorm kernel itself
  }
  void layernorm(size_t x_off, size_t out_off, int nrows, int C, const Vec& g, const Vec& b, const Vec* pe,
                 int rows_per_frame) {
    if (dry || failed) return;
    const void* x = p(x_off);
    void* o = p(out_off);
    const float* gp = wv(g);
    const float* bp = wv(b);
    const float* pep = pe ? wv(*pe) : nullptr;
    const int frames = h->F, dt = h->dt;
    if ((C / 8 + 31) / 32 > 5) return fail("layernorm: C too large");
    push([=](cudaStream_t s) { ln_run(dt, x, o, gp, bp, nrows, C, 1e-5f, pep, rows_per_frame, frames, s); }, "layernorm",
         0.0, 2.0 * nrows * C * 2.0);
  }
  void attention(AttnDesc d) {
    if (dry || failed) return;
    d.dt = h->dt;
    const double flops = 4.0 * d.batch * d.heads * (double)d.S_q * d.S_kv * d.d;
    const double bytes = 2.0 * d.batch * d.heads * d.d * (2.0 * d.S_q + 2.0 * d.S_kv);
    const char* kind = d.S_q == d.S_kv ? "attn_self" : "attn_cross";
    if (h->simple) {
      push([d](cudaStream_t s) {
        attn_simple_launch(d, s);
        g_launches++;
      }, kind, flops, bytes);
      return;
    }
    AttnLaunch l;
    std::string e;
    if (!attn_prepare(d, &l, &e)) return fail(e);
    push([l](cudaStream_t s) {
      attn_launch(l, s);
      g_launches++;
    }, kind, flops, bytes, d.S_q, d.S_kv, d.d);
  }
  void tap(const std::string& name, const Act& a) {
    if (!h->taps_enabled) return;
    const size_t bytes = abytes(a);
    const size_t off = alloc(bytes);  // never released
    if (dry) return;
    h->taps[name] = TapInfo{off, rows(a), a.C};
    void* dst = p(off);
    const void* src = p(a.off);
    push([=](cudaStream_t s) { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s); });
  }

  // ---------------- modules ----------------
  // ResnetBlock3D.forward (resnet.py:182-212); in1 = skip tensor concatenated on the channel dim (unet_blocks.py:644,754)
  Act resnet(const Act& in0, const Act* in1, const ResW& r) {
    const int cin = in0.C + (in1 ? in1->C : 0);
    if (cin != r.cin) fail("resnet: channel mismatch");
    if (!r.shortcut && in1) fail("resnet: concat input without shortcut");
    const int NI = h->B * h->F, M = rows(in0);
    Act n = new_act(cin, in0.H, in0.W);
    groupnorm(in0, in1, r.n1g, r.n1b, h->cfg.norm_eps, false, true, n.off);
    Act h1 = new_act(r.cout, in0.H, in0.W);
    h1.gn = new_gn(r.cout, in0.H, in0.W);
    {
      GemmDesc d;
      memset(&d, 0, sizeof d);
      if (h1.gn != NO_GN) {
        d.gn_acc = gnp(h1.gn);
        d.gn_hw = in0.H * in0.W;
      }
      d.M = M;
      d.N = r.cout;
      d.nseg = 1;
      d.seg[0] = ASeg{SEG_CONV3, p(n.off), cin, cin, in0.H, in0.W, NI};
      d.w = wm(r.c1);
      d.Ktot = r.c1.cols;
      d.w_rows = r.cout;
      d.Ho = in0.H;
      d.Wo = in0.W;
      d.NI = NI;
      d.out = p(h1.off);
      d.ldo = r.cout;
      d.bias = wv(h->bias_eff_all) + r.temb_row;  // conv1.bias + time_emb_proj(silu(emb)) for this step
      gemm(d);
    }
    free_act(n);
    Act n2 = new_act(r.cout, in0.H, in0.W);
    groupnorm(h1, nullptr, r.n2g, r.n2b, h->cfg.norm_eps, false, true, n2.off);
    free_act(h1);
    Act out = new_act(r.cout, in0.H, in0.W);
    out.gn = new_gn(r.cout, in0.H, in0.W);
    {
      GemmDesc d;
      memset(&d, 0, sizeof d);
      if (out.gn != NO_GN) {
        d.gn_acc = gnp(out.gn);
        d.gn_hw = in0.H * in0.W;
      }
      d.M = M;
      d.N = r.cout;
      d.nseg = 1;
      d.seg[0] = ASeg{SEG_CONV3, p(n2.off), r.cout, r.cout, in0.H, in0.W, NI};
      if (r.shortcut) {  // 1x1 shortcut over the (virtual) concat accumulates into the same tile
        d.seg[d.nseg++] = ASeg{SEG_PLAIN, p(in0.off), in0.C, in0.C, 0, 0, 0};
        if (in1) d.seg[d.nseg++] = ASeg{SEG_PLAIN, p(in1->off), in1->C, in1->C, 0, 0, 0};
      } else {
        d.res = p(in0.off);
        d.ldr = in0.C;
      }
      d.w = wm(r.c2);
      d.Ktot = r.c2.cols;
      d.w_rows = r.cout;
      d.Ho = in0.H;
      d.Wo = in0.W;
      d.NI = NI;
      d.out = p(out.off);
      d.ldo = r.cout;
      d.bias = wv(r.c2beff);
      gemm(d);
    }
    free_act(n2);
    return out;
  }

  // Transformer3DModel + BasicTransformerBlock (attention.py:318-365, 479-526); x updated in place
  void transformer(Act& x, const TfW& t, size_t kv_off) {
    const int C = t.C, M = rows(x), HW = x.H * x.W, NI = h->B * h->F;
    const int heads = h->cfg.attention_heads, d = C / heads;
    Act n = new_act(C, x.H, x.W);
    groupnorm(x, nullptr, t.ng, t.nb, 1e-6f, true, false, n.off);
    const bool fold = h->ln_fold && !h->simple;
    // row statistics of y for the folded LayerNorms: [column parts][M] float2, rewritten by every producer of y
    const size_t st_bytes = (size_t)gemm_stats_parts(C, M) * M * sizeof(float2);
    const size_t st = fold ? alloc(st_bytes) : 0;
    Act y = new_act(C, x.H, x.W);
    linear(n.off, M, C, t.pi, &t.pib, y.off, C, nullptr, 0, nullptr, st, fold);
    free_act(n);
    Act tmp = new_act(C, x.H, x.W);
    // self-attention
    Act qkv = new_act(3 * C, x.H, x.W);
    if (fold) {
      linear(y.off, M, C, t.a1.qkv, nullptr, qkv.off, 3 * C, nullptr, 0, &t.a1.qkv_ln, st);
    } else {
      layernorm(y.off, tmp.off, M, C, t.ln1g, t.ln1b, nullptr, 1);
      linear(tmp.off, M, C, t.a1.qkv, nullptr, qkv.off, 3 * C, nullptr);
    }
    {
      AttnDesc a;
      memset(&a, 0, sizeof a);
      a.q = p(qkv.off);
      a.ldq = 3 * C;
      a.k = reinterpret_cast<const char*>(p(qkv.off)) + (size_t)C * 2;
      a.v = reinterpret_cast<const char*>(p(qkv.off)) + (size_t)2 * C * 2;
      a.ldkv = 3 * C;
      a.S_q = HW;
      a.S_kv = HW;
      a.heads = heads;
      a.d = d;
      a.batch = NI;
      a.out = p(tmp.off);
      a.ldo = C;
      attention(a);
    }
    free_act(qkv);
    linear(tmp.off, M, C, t.a1.out, &t.a1.outb, y.off, C, &y.off, 0, nullptr, st, fold);
    // cross-attention to the fused context
    Act q = new_act(C, x.H, x.W);
    if (fold) {
      linear(y.off, M, C, t.a2.q, nullptr, q.off, C, nullptr, 0, &t.a2.q_ln, st);
    } else {
      layernorm(y.off, tmp.off, M, C, t.ln2g, t.ln2b, nullptr, 1);
      linear(tmp.off, M, C, t.a2.q, nullptr, q.off, C, nullptr);
    }
    {
      AttnDesc a;
      memset(&a, 0, sizeof a);
      a.q = p(q.off);
      a.ldq = C;
      a.k = p(kv_off);
      a.v = reinterpret_cast<const char*>(p(kv_off)) + (size_t)C * 2;
      a.ldkv = 2 * C;
      a.S_q = HW;
      a.S_kv = h->L;
      a.heads = heads;
      a.d = d;
      a.batch = NI;
      a.out = p(tmp.off);
      a.ldo = C;
      attention(a);
    }
    free_act(q);
    linear(tmp.off, M, C, t.a2.out, &t.a2.outb, y.off, C, &y.off, 0, nullptr, st, fold);
    // GEGLU feed-forward
    if (ffn_fused_ok(C, fold)) {
      ffn_fused(y.off, M, t.ff1_ln, st, t.ff2, t.ff2b);
    } else {
      Act g = new_act(4 * C, x.H, x.W);
      if (fold) {
        linear(y.off, M, C, t.ff1, &t.ff1b, g.off, 4 * C, nullptr, 1, &t.ff1_ln, st);
      } else {
        layernorm(y.off, tmp.off, M, C, t.ln3g, t.ln3b, nullptr, 1);
        linear(tmp.off, M, C, t.ff1, &t.ff1b, g.off, 4 * C, nullptr, 1);
      }
      if (po_fold_ok(C)) {
        free_act(tmp);
        x.gn = new_gn(C, x.H, x.W);  // x is rewritten in place: its statistics are new
        proj_out_folded(y.off, g.off, M, C, t.pof, t.pofb, x.off, x.gn, HW);
        free_act(g);
        free_act(y);
        if (fold) release(st, st_bytes);
        return;
      }
      linear(g.off, M, 4 * C, t.ff2, &t.ff2b, y.off, C, &y.off);
      free_act(g);
    }
    free_act(tmp);
    x.gn = new_gn(C, x.H, x.W);  // x is rewritten in place: its statistics are new
    linear(y.off, M, C, t.po, &t.pob, x.off, C, &x.off, 0, nullptr, 0, false, 1, x.gn, HW);
    free_act(y);
    if (fold) release(st, st_bytes);
  }

  // VanillaTemporalModule (motion_module.py:87-93,147-182,234-246,294-354); x updated in place
  void motion(Act& x, const MoW& m) {
    const int C = m.C, M = rows(x), HW = x.H * x.W;
    const int heads = h->cfg.motion_heads, d = C / heads;
    Act n = new_act(C, x.H, x.W);
    groupnorm(x, nullptr, m.ng, m.nb, 1e-6f, true, false, n.off);
    const bool fold = h->ln_fold && !h->simple;
    const size_t st_bytes = (size_t)gemm_stats_parts(C, M) * M * sizeof(float2);
    const size_t st = fold ? alloc(st_bytes) : 0;
    Act y = new_act(C, x.H, x.W);
    linear(n.off, M, C, m.pi, &m.pib, y.off, C, nullptr, 0, nullptr, st, fold);
    free_act(n);
    Act tmp = new_act(C, x.H, x.W);
    for (int i = 0; i < h->cfg.motion_attn_blocks; ++i) {
      Act qkv = new_act(3 * C, x.H, x.W);
      if (fold) {  // LayerNorm + positional encoding folded into the QKV projection (c is per frame)
        linear(y.off, M, C, m.att[i].qkv, nullptr, qkv.off, 3 * C, nullptr, 0, &m.att[i].qkv_ln, st, false, HW);
      } else {
        layernorm(y.off, tmp.off, M, C, m.lng[i], m.lnb[i], &m.pe[i], HW);
        linear(tmp.off, M, C, m.att[i].qkv, nullptr, qkv.off, 3 * C, nullptr);
      }
      if (!dry && !failed) {
        const void* qp = p(qkv.off);
        void* op = p(tmp.off);
        const int dt = h->dt, B = h->B, F = h->F;
        push([=](cudaStream_t s) {
          temporal_attn_launch(dt, qp, op, B, F, HW, heads, d, s);
          g_launches++;
        }, "attn_temporal", 4.0 * B * HW * heads * (double)F * F * d, 2.0 * M * 4.0 * C);
      }
      free_act(qkv);
      linear(tmp.off, M, C, m.att[i].out, &m.att[i].outb, y.off, C, &y.off, 0, nullptr, st, fold);
    }
    if (ffn_fused_ok(C, fold)) {
      ffn_fused(y.off, M, m.ff1_ln, st, m.ff2, m.ff2b);
    } else {
      Act g = new_act(4 * C, x.H, x.W);
      if (fold) {
        linear(y.off, M, C, m.ff1, &m.ff1b, g.off, 4 * C, nullptr, 1, &m.ff1_ln, st);
      } else {
        layernorm(y.off, tmp.off, M, C, m.ffng, m.ffnb, nullptr, 1);
        linear(tmp.off, M, C, m.ff1, &m.ff1b, g.off, 4 * C, nullptr, 1);
      }
      if (po_fold_ok(C)) {
        free_act(tmp);
        x.gn = new_gn(C, x.H, x.W);
        proj_out_folded(y.off, g.off, M, C, m.pof, m.pofb, x.off, x.gn, HW);
        free_act(g);
        free_act(y);
        if (fold) release(st, st_bytes);
        return;
      }
      linear(g.off, M, 4 * C, m.ff2, &m.ff2b, y.off, C, &y.off);
      free_act(g);
    }
    free_act(tmp);
    x.gn = new_gn(C, x.H, x.W);
    linear(y.off, M, C, m.po, &m.pob, x.off, C, &x.off, 0, nullptr, 0, false, 1, x.gn, HW);
    free_act(y);
    if (fold) release(st, st_bytes);
  }

  Act conv_sampler(const Act& x, const Mat& w, const Vec& b, int stride, int Ho, int Wo) {
    Act out = new_act(w.rows, Ho, Wo);
    out.gn = new_gn(w.rows, Ho, Wo);
    GemmDesc d;
    memset(&d, 0, sizeof d);
    if (out.gn != NO_GN) {
      d.gn_acc = gnp(out.gn);
      d.gn_hw = Ho * Wo;
    }
    d.M = rows(out);
    d.N = w.rows;
    d.nseg = 1;
    d.seg[0] = ASeg{stride == 2 ? SEG_CONV3S2 : SEG_CONV3, p(x.off), x.C, x.C, x.H, x.W, h->B * h->F};
    d.w = wm(w);
    d.Ktot = w.cols;
    d.w_rows = w.rows;
    d.Ho = Ho;
    d.Wo = Wo;
    d.NI = h->B * h->F;
    d.out = p(out.off);
    d.ldo = w.rows;
    d.bias = wv(b);
    gemm(d);
    return out;
  }
};

// walk the network once; records ops (unless dry) and returns the peak arena size
static int plan_network(rcdm_unet_impl* h, bool dry, size_t* peak) {
  const rcdm_unet_config& c = h->cfg;
  Planner P{h, dry, nullptr};
  const int NI = h->B * h->F;
  // ---- persistent regions
  const size_t gn_bytes = gn_scratch_bytes(NI, c.norm_num_groups);
  P.gn_scratch = P.alloc(gn_bytes);
  const size_t gn_region_bytes = dry ? 0 : h->gn_acc_bytes;  // measured by the dry pass
  P.gn_region = gn_region_bytes ? P.alloc(gn_region_bytes) : 0;
  const int c0 = c.block_out_channels[0];
  const size_t e1 = P.alloc((size_t)h->temb_dim * 4), e2 = P.alloc((size_t)h->temb_dim * 4);
  const size_t ctx16 = P.alloc((size_t)NI * h->L * c.cross_attention_dim * 2);
  // ---- ctx ops
  P.ops = &h->ctx_ops;
  if (!dry) {
    void* dst = P.p(ctx16);
    const size_t n = (size_t)NI * h->L * c.cross_attention_dim;
    const int dt = h->dt;
    void* gscr = P.p(P.gn_scratch);
    P.push([=](cudaStream_t s) {
      cudaMemsetAsync(gscr, 0, gn_bytes, s);
      if (dt == DT_F16)
        cast_rows_kernel<__half><<<grid_for(n, 256), 256, 0, s>>>(h->cur_ctx, h->cur_ctx_dt,
                                                                  reinterpret_cast<__half*>(dst), n);
      else
        cast_rows_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, s>>>(h->cur_ctx, h->cur_ctx_dt,
                                                                         reinterpret_cast<__nv_bfloat16*>(dst), n);
      g_launches++;
    });
  }
  // cross-attention K/V of every spatial transformer: step-invariant, computed by ctx_ops BEFORE any step op runs,
  // so these buffers are persistent and must be carved out before any temporary is allocated.
  std::map<const TfW*, size_t> kv_of;
  {
    std::vector<const TfW*> tfs;
    for (auto& blk : h->down)
      for (auto& l : blk.layers)
        if (l.has_tf) tfs.push_back(&l.tf);
    tfs.push_back(&h->mid_tf);
    for (auto& blk : h->up)
      for (auto& l : blk.layers)
        if (l.has_tf) tfs.push_back(&l.tf);
    for (const TfW* t : tfs) {
      const size_t off = P.alloc((size_t)NI * h->L * 2 * t->C * 2);
      kv_of[t] = off;
      P.linear(ctx16, NI * h->L, c.cross_attention_dim, t->a2.kv, nullptr, off, 2 * t->C, nullptr);
    }
  }
  auto plan_kv = [&](const TfW& t) { return kv_of.at(&t); };

  // ---- step ops
  P.ops = &h->step_ops;
  if (!dry && gn_region_bytes) {  // GroupNorm chunk accumulators of every normalised tensor: zero at the start of a step
    void* reg = P.p(P.gn_region);
    P.push([=](cudaStream_t s) { cudaMemsetAsync(reg, 0, gn_region_bytes, s); });
  }
  if (!dry) {
    float* e1p = reinterpret_cast<float*>(P.p(e1));
    float* e2p = reinterpret_cast<float*>(P.p(e2));
    const int dt = h->dt, td = h->temb_dim, tr = h->temb_rows;
    const void* l1w = P.wm(h->l1w);
    const void* l2w = P.wm(h->l2w);
    const void* tw = P.wm(h->temb_all);
    const float* l1b = P.wv(h->l1b);
    const float* l2b = P.wv(h->l2b);
    const float* sb = P.wv(h->temb_static_b);
    float* beff = const_cast<float*>(P.wv(h->bias_eff_all));
    const int flip = c.flip_sin_to_cos;
    const float fs = c.freq_shift;
    P.push([=](cudaStream_t s) {
      // unet.py:367-389 + every resnet's time_emb_proj (resnet.py:190-191) in three launches
      if (dt == DT_F16) {
        temb_linear1_kernel<__half><<<(td * 32 + 255) / 256, 256, 0, s>>>(h->cur_t_dev, h->cur_t_host,
                                                                          reinterpret_cast<const __half*>(l1w), l1b, e1p,
                                                                          c0, td, flip, fs);
        gemv_kernel<__half><<<(td * 32 + 255) / 256, 256, 0, s>>>(e1p, reinterpret_cast<const __half*>(l2w), l2b,
                                                                  nullptr, e2p, td, td, 1);
        gemv_kernel<__half><<<(tr * 32 + 255) / 256, 256, 0, s>>>(e2p, reinterpret_cast<const __half*>(tw), sb, nullptr,
                                                                  beff, td, tr, 0);
      } else {
        temb_linear1_kernel<__nv_bfloat16><<<(td * 32 + 255) / 256, 256, 0, s>>>(
            h->cur_t_dev, h->cur_t_host, reinterpret_cast<const __nv_bfloat16*>(l1w), l1b, e1p, c0, td, flip, fs);
        gemv_kernel<__nv_bfloat16><<<(td * 32 + 255) / 256, 256, 0, s>>>(
            e1p, reinterpret_cast<const __nv_bfloat16*>(l2w), l2b, nullptr, e2p, td, td, 1);
        gemv_kernel<__nv_bfloat16><<<(tr * 32 + 255) / 256, 256, 0, s>>>(
            e2p, reinterpret_cast<const __nv_bfloat16*>(tw), sb, nullptr, beff, td, tr, 0);
      }
      g_launches += 3;
    });
  }
  // conv_in via im2col (9 input channels) + tensor-core GEMM
  Act x = P.new_act(c0, h->H, h->W);
  {
    const int M = P.rows(x), kpad = h->conv_in_kpad;
    const size_t a_off = P.alloc((size_t)M * kpad * 2);
    if (!dry) {
      void* A = P.p(a_off);
      const int dt = h->dt, B = h->B, Cin = c.in_channels, F = h->F, H = h->H, W = h->W;
      P.push([=](cudaStream_t s) {
        const size_t total = (size_t)M * kpad;
        if (dt == DT_F16)
          im2col_in_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(h->cur_sample, h->cur_sample_dt,
                                                                        reinterpret_cast<__half*>(A), B, Cin, F, H, W,
                                                                        kpad);
        else
          im2col_in_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(
              h->cur_sample, h->cur_sample_dt, reinterpret_cast<__nv_bfloat16*>(A), B, Cin, F, H, W, kpad);
        g_launches++;
      });
    }
    x.gn = P.new_gn(c0, h->H, h->W);
    P.linear(a_off, M, kpad, h->conv_in_w, &h->conv_in_b, x.off, c0, nullptr, 0, nullptr, 0, false, 1, x.gn, h->H * h->W);
    P.release(a_off, (size_t)M * kpad * 2);
  }
  P.tap("conv_in", x);
  std::vector<Act> skips{x};
  auto in_skips = [&](const Act& a) {
    for (auto& s : skips)
      if (s.off == a.off) return true;
    return false;
  };
  for (int i = 0; i < c.num_blocks; ++i) {
    BlockW& blk = h->down[i];
    const std::string bp = "down_blocks." + std::to_string(i);
    for (size_t j = 0; j < blk.layers.size(); ++j) {
      LayerW& l = blk.layers[j];
      Act r = P.resnet(x, nullptr, l.res);
      if (!in_skips(x)) P.free_act(x);
      x = r;
      P.tap(bp + ".resnets." + std::to_string(j), x);
      if (l.has_tf) {
        P.transformer(x, l.tf, plan_kv(l.tf));
        P.tap(bp + ".attentions." + std::to_string(j), x);
      }
      if (l.has_mo) {
        P.motion(x, l.mo);
        P.tap(bp + ".motion_modules." + std::to_string(j), x);
      }
      skips.push_back(x);
    }
    if (blk.sampler) {
      if (x.H % 2 || x.W % 2) {
        P.fail("downsample needs even spatial dims");
        break;
      }
      x = P.conv_sampler(x, blk.sw, blk.sb, 2, x.H / 2, x.W / 2);
      P.tap(bp + ".downsamplers.0", x);
      skips.push_back(x);
    }
  }
  {  // mid block (unet_blocks.py:272-280)
    Act r0 = P.resnet(x, nullptr, h->mid_r0);
    P.transformer(r0, h->mid_tf, plan_kv(h->mid_tf));
    if (h->mid_has_mo) P.motion(r0, h->mid_mo);
    Act r1 = P.resnet(r0, nullptr, h->mid_r1);
    P.free_act(r0);
    x = r1;  // previous x is the last skip: still owned by `skips`
    P.tap("mid_block", x);
  }
  for (int i = 0; i < c.num_blocks && !P.failed; ++i) {
    BlockW& blk = h->up[i];
    const std::string bp = "up_blocks." + std::to_string(i);
    for (size_t j = 0; j < blk.layers.size(); ++j) {
      LayerW& l = blk.layers[j];
      if (skips.empty()) {
        P.fail("skip stack underflow");
        break;
      }
      Act sk = skips.back();
      skips.pop_back();
      if (sk.H != x.H || sk.W != x.W) {
        P.fail("skip / hidden spatial mismatch");
        break;
      }
      Act r = P.resnet(x, &sk, l.res);
      P.free_act(x);
      P.free_act(sk);
      x = r;
      P.tap(bp + ".resnets." + std::to_string(j), x);
      if (l.has_tf) {
        P.transformer(x, l.tf, plan_kv(l.tf));
        P.tap(bp + ".attentions." + std::to_string(j), x);
      }
      if (l.has_mo) {
        P.motion(x, l.mo);
        P.tap(bp + ".motion_modules." + std::to_string(j), x);
      }
    }
    if (blk.sampler && !h->simple) {
      // Upsample3D (resnet.py:46-80): nearest 2x folded into the conv - four launches (one per output parity class) of a
      // 2x2 conv on the un-upsampled activation, K = 4 C; the 4x tensor is never materialised
      Act o = P.new_act(blk.C, 2 * x.H, 2 * x.W);
      o.gn = P.new_gn(blk.C, x.H, x.W);  // statistics per image from the four class launches (class grid = x's grid)
      for (int cls = 0; cls < 4; ++cls) {
        GemmDesc d;
        memset(&d, 0, sizeof d);
        if (o.gn != NO_GN) {
          d.gn_acc = P.gnp(o.gn);
          d.gn_hw = x.H * x.W;
        }
        d.M = P.rows(x);
        d.N = blk.C;
        d.nseg = 1;
        d.seg[0] = ASeg{SEG_UP2, P.p(x.off), x.C, x.C, x.H, x.W, NI};
        d.w = h->arena + blk.swf.off + (size_t)cls * blk.C * 4 * x.C * 2;
        d.Ktot = 4 * x.C;
        d.w_rows = blk.C;
        d.Ho = x.H;
        d.Wo = x.W;
        d.NI = NI;
        d.out = P.p(o.off);
        d.ldo = blk.C;
        d.bias = P.wv(blk.sb);
        d.up_py = cls >> 1;
        d.up_px = cls & 1;
        P.gemm(d);
      }
      P.free_act(x);
      x = o;
      P.tap(bp + ".upsamplers.0", x);
    } else if (blk.sampler) {  // debug (CUDA-core) plan: materialised nearest 2x, then the conv
      Act u = P.new_act(x.C, 2 * x.H, 2 * x.W);
      if (!dry) {
        const void* src = P.p(x.off);
        void* dst = P.p(u.off);
        const int dt = h->dt, N = NI, H = x.H, W = x.W, C = x.C;
        P.push([=](cudaStream_t s) {
          const size_t total = (size_t)N * 4 * H * W * (C / 8);
          if (dt == DT_F16)
            upsample2x_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const __half*>(src),
                                                                            reinterpret_cast<__half*>(dst), N, H, W, C);
          else
            upsample2x_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(
                reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(dst), N, H, W, C);
          g_launches++;
        });
      }
      P.free_act(x);
      x = P.conv_sampler(u, blk.sw, blk.sb, 1, u.H, u.W);
      P.free_act(u);
      P.tap(bp + ".upsamplers.0", x);
    }
  }
  if (!P.failed) {  // conv_norm_out -> SiLU -> conv_out (unet.py:455-457), then back to NCFHW
    Act n = P.new_act(x.C, x.H, x.W);
    P.groupnorm(x, nullptr, h->cno_g, h->cno_b, c.norm_eps, false, true, n.off);
    P.free_act(x);
    Act o = P.new_act(8, x.H, x.W);  // only out_channels (<= 8) columns are written
    if (c.out_channels > 8) P.fail("out_channels > 8 unsupported");
    {
      GemmDesc d;
      memset(&d, 0, sizeof d);
      d.M = P.rows(n);
      d.N = c.out_channels;
      d.nseg = 1;
      d.seg[0] = ASeg{SEG_CONV3, P.p(n.off), n.C, n.C, n.H, n.W, NI};
      d.w = P.wm(h->conv_out_w);
      d.Ktot = h->conv_out_w.cols;
      d.w_rows = c.out_channels;
      d.Ho = n.H;
      d.Wo = n.W;
      d.NI = NI;
      d.out = P.p(o.off);
      d.ldo = c.out_channels;
      d.bias = P.wv(h->conv_out_b);
      P.gemm(d);
    }
    P.free_act(n);
    if (!dry) {
      const void* tok = P.p(o.off);
      const int dt = h->dt, B = h->B, C = c.out_channels, F = h->F, HW = h->H * h->W;
      P.push([=](cudaStream_t s) {
        const size_t total = (size_t)B * C * F * HW;
        if (dt == DT_F16)
          tokens_to_ncfhw_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const __half*>(tok),
                                                                              h->cur_out, h->cur_out_dt, B, C, F, HW);
        else
          tokens_to_ncfhw_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(
              reinterpret_cast<const __nv_bfloat16*>(tok), h->cur_out, h->cur_out_dt, B, C, F, HW);
        g_launches++;
      });
    }
    P.free_act(o);
  }
  if (P.failed) return set_err("plan: " + P.err);
  if (dry) h->gn_acc_bytes = P.gn_used;
  else if (P.gn_used != h->gn_acc_bytes) return set_err("plan: GroupNorm accumulator region changed between passes");
  *peak = P.peak + (dry ? h->gn_acc_bytes + 1024 : 0);
  return 0;
}

int unet_create(const rcdm_unet_config* cfg, rcdm_unet** out) {
  if (!cfg || !out) return set_err("null argument");
  rcdm_unet* h = new rcdm_unet();
  h->cfg = *cfg;
  // debug switches (rcdm_unet_set_option before rcdm_unet_prepare): simple = 0, autotune = 0 (measured: no gain over the
  // heuristics at the 512x512 shapes), ln_fold = 1, gn_stats = library option at creation time
  h->gn_stats = opt(OPT_GN_STATS);
  h->ffn_pack64 = opt(OPT_FFN_FUSED) ? 1 : 0;
  if (build_model(h)) {
    delete h;
    return 1;
  }
  *out = h;
  return 0;
}

static void release_plan(rcdm_unet_impl* h) {
  h->ctx_ops.clear();
  h->step_ops.clear();
  h->step_meta.clear();
  h->taps.clear();
  if (h->graph_exec) {
    cudaGraphExecDestroy(h->graph_exec);
    h->graph_exec = nullptr;
  }
  h->graph_key.clear();
  if (h->ws) {
    cudaFree(h->ws);
    h->ws = nullptr;
  }
  h->planned = false;
}

void unet_destroy(rcdm_unet* h) {
  if (!h) return;
  release_plan(h);
  if (h->arena) cudaFree(h->arena);
  sk_workspace_free(&h->sk);
  if (h->pack_jobs) cudaFree(h->pack_jobs);
  if (h->loop_buf) cudaFree(h->loop_buf);
  if (h->loop_stream) cudaStreamDestroy(h->loop_stream);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  if (h->ev_out) cudaEventDestroy(h->ev_out);
  delete h;
}

static int ensure_arena(rcdm_unet_impl* h) {
  if (h->arena) return 0;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    return set_err("no CUDA device: librcdm_b200 has no CPU fallback");
  }
  CUDA_OK(cudaMalloc(&h->arena, h->arena_bytes));
  CUDA_OK(cudaMemset(h->arena, 0, h->arena_bytes));
  std::string e;
  if (!gemm_setup_attributes(&e) || !attn_setup_attributes(&e) || !gn_setup_attributes(&e) || !ffn_setup_attributes(&e)) return set_err(e);
  if (!sk_workspace_alloc(&h->sk, &e)) return set_err(e);
  return 0;
}

int unet_load_weight(rcdm_unet* h, const char* name, const void* data, int dtype, const int64_t* dims, int ndim,
                     void* stream) {
  if (!h || !name || !data) return set_err("null argument");
  auto it = h->slot_index.find(name);
  if (it == h->slot_index.end()) return set_err(std::string("unexpected key in state_dict: ") + name);
  Slot& s = h->slots[it->second];
  if (ndim != s.ndim) return set_err(std::string("size mismatch for ") + name);
  for (int i = 0; i < ndim; ++i)
    if (dims[i] != s.dims[i]) return set_err(std::string("size mismatch for ") + name);
  if (dtype < 0 || dtype > 2) return set_err("bad dtype");
  if (s.kind == SLOT_IGNORE) {
    s.loaded = true;
    return 0;
  }
  if (ensure_arena(h)) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (h->dt == DT_F16) pack_dispatch<__half>(h, s, data, dtype, st);
  else pack_dispatch<__nv_bfloat16>(h, s, data, dtype, st);
  CUDA_OK(cudaGetLastError());
  s.loaded = true;
  h->dirty = true;
  return 0;
}

// The whole state dict in one launch (a model load used to be 1 286 one-tensor launches).
int unet_load_weights(rcdm_unet* h, int count, const char* const* names, const void* const* data, const int* dtypes,
                      const int64_t* dims, const int* ndims, void* stream) {
  if (!h || !names || !data || !dtypes || !dims || !ndims || count < 0) return set_err("null argument");
  std::vector<PackJob> jobs;
  std::vector<int> touched;
  jobs.reserve(count);
  long long block = 0;
  for (int i = 0; i < count; ++i) {
    if (!names[i] || !data[i]) return set_err("null state-dict entry");
    auto it = h->slot_index.find(names[i]);
    if (it == h->slot_index.end()) return set_err(std::string("unexpected key in state_dict: ") + names[i]);
    const Slot& s = h->slots[it->second];
    if (ndims[i] != s.ndim) return set_err(std::string("size mismatch for ") + names[i]);
    for (int k = 0; k < s.ndim; ++k)
      if (dims[(size_t)i * 4 + k] != s.dims[k]) return set_err(std::string("size mismatch for ") + names[i]);
    if (dtypes[i] < 0 || dtypes[i] > 2) return set_err("bad dtype");
    touched.push_back(it->second);
    if (s.kind == SLOT_IGNORE) continue;
    PackJob j;
    memset(&j, 0, sizeof j);
    j.src = data[i];
    j.src_dt = dtypes[i];
    j.N = (int)s.dims[0];
    j.K = 1;
    for (int k = 1; k < s.ndim; ++k) j.K *= (int)s.dims[k];
    if (s.kind == SLOT_VEC) {
      j.N *= j.K;
      j.K = 1;
      j.kind = 2;
    } else {
      j.kind = s.kind == SLOT_CONV3 ? 1 : 0;
    }
    j.ldd = s.ldd;
    j.col_off = s.col_off;
    j.row_off = s.row_off;
    j.Cin = s.cin;
    j.geglu_bn = s.geglu_bn;
    j.dst = nullptr;  // resolved below (the arena may not exist yet)
    const size_t total = (size_t)j.N * j.K;
    size_t nb = (total + 256 * 16 - 1) / (256 * 16);
    j.nblocks = (int)(nb < 1 ? 1 : nb > 4096 ? 4096 : nb);
    j.block0 = (int)block;
    block += j.nblocks;
    jobs.push_back(j);
  }
  if (ensure_arena(h)) return 1;
  {
    size_t ji = 0;
    for (int i = 0; i < count; ++i) {
      const Slot& s = h->slots[touched[i]];
      if (s.kind == SLOT_IGNORE) continue;
      jobs[ji++].dst = h->arena + s.dst;
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!jobs.empty()) {
    const size_t bytes = jobs.size() * sizeof(PackJob);
    if (h->pack_jobs_bytes < bytes) {
      if (h->pack_jobs) CUDA_OK(cudaFree(h->pack_jobs));
      h->pack_jobs = nullptr;
      CUDA_OK(cudaMalloc(&h->pack_jobs, bytes));
      h->pack_jobs_bytes = bytes;
    }
    // pageable source: the copy is staged before the call returns, so the host vector may go out of scope
    CUDA_OK(cudaMemcpyAsync(h->pack_jobs, jobs.data(), bytes, cudaMemcpyHostToDevice, st));
    const PackJob* jd = reinterpret_cast<const PackJob*>(h->pack_jobs);
    if (h->dt == DT_F16) pack_many_kernel<__half><<<(unsigned)block, 256, 0, st>>>(jd, (int)jobs.size());
    else pack_many_kernel<__nv_bfloat16><<<(unsigned)block, 256, 0, st>>>(jd, (int)jobs.size());
    g_launches++;
    CUDA_OK(cudaGetLastError());
  }
  for (int idx : touched) h->slots[idx].loaded = true;
  h->dirty = true;
  return 0;
}

int unet_prepare(rcdm_unet* h, int batch, int frames, int height, int width, int ctx_len) {
  if (!h) return set_err("null handle");
  if (batch < 1 || frames < 1 || frames > 5 || height < 1 || width < 1 || ctx_len < 1)
    return set_err("prepare: bad problem size (frames must be 1..5)");
  if (h->planned && h->B == batch && h->F == frames && h->H == height && h->W == width && h->L == ctx_len) return 0;
  if (ensure_arena(h)) return 1;
  release_plan(h);
  h->B = batch;
  h->F = frames;
  h->H = height;
  h->W = width;
  h->L = ctx_len;
  size_t peak = 0;
  if (plan_network(h, true, &peak)) return 1;
  h->ws_bytes = peak + 4096;
  CUDA_OK(cudaMalloc(&h->ws, h->ws_bytes));
  CUDA_OK(cudaMemset(h->ws, 0, h->ws_bytes));
  if (plan_network(h, false, &peak)) {
    release_plan(h);
    return 1;
  }
  h->planned = true;
  return 0;
}

int unet_run(rcdm_unet* h, bool run_ctx, bool run_step, cudaStream_t st) {
  if (h->dirty) finalize_weights(h, st);
  if (run_ctx)
    for (auto& op : h->ctx_ops) op(st);
  if (run_step)
    for (auto& op : h->step_ops) op(st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(std::string("kernel launch failed: ") + cudaGetErrorString(e));
  return 0;
}

}  // namespace rcdm
