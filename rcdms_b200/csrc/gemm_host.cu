// Host side of the tcgen05 GEMM / implicit-GEMM conv: tensor-map construction, tile selection, launch.
// Also holds the CUDA-core "simple" kernel with identical semantics, used only for bisecting parity
// failures (RCDM_SIMPLE=1); it is never the benchmarked path.
#include <cudaTypedefs.h>

#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>

#include "launch.h"

namespace rcdm {

// ------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn(std::string* err) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  static std::string once_err;
  std::call_once(once, [&]() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      once_err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
      (void)cudaGetLastError();
      return;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  if (!fn && err) *err = once_err;
  return fn;
}

// L2 promotion of TMA loads (experiment switch; 256 B measured best / equal on the UNet step)
#ifndef RCDM_TMA_L2_PROMOTION
#define RCDM_TMA_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif
bool encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, bool swizzle128, std::string* err) {
  return encode_tmap_sw(out, base, rank, dims, strides_bytes, box, swizzle128 ? 128 : 0, err);
}
// swizzle_bytes: 0 (none) | 32 | 64 | 128
bool encode_tmap_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes, std::string* err) {
  auto fn = get_encode_fn(err);
  if (!fn) return false;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i < rank - 1) gstr[i] = strides_bytes[i];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    if (err) *err = "TMA base address not 16-byte aligned";
    return false;
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  RCDM_TMA_L2_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[256];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
               (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
               (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
               (unsigned long long)(rank > 4 ? gdim[4] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
               rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0);
      *err = buf;
    }
    return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// GEMM prepare / launch
// ------------------------------------------------------------------------------------------
static std::atomic<int> g_opts[OPT_COUNT] = {{0}, {24}, {1}, {1}, {1}, {1}, {1}, {40}, {1}, {1}, {1}, {1}, {0}};
static const char* const g_opt_names[OPT_COUNT] = {"pdl", "sk_min", "gemm_pair", "masked_attn_mma",
                                                    "temporal_wide", "temporal_wide_all", "temporal_tiled",
                                                    "temporal_smem_kb", "gn_fused", "ln_wide", "gn_stats", "attn_short_kv",
                                                    "ffn_fused"};
int opt(int id) { return g_opts[id].load(std::memory_order_relaxed); }
int opt_set(const char* name, int value, int* previous) {
  for (int i = 0; i < OPT_COUNT; ++i)
    if (name && !strcmp(name, g_opt_names[i])) {
      const int prev = g_opts[i].exchange(value);
      if (previous) *previous = prev;
      return 0;
    }
  return 1;
}

int pdl_mode() { return opt(OPT_PDL); }
bool pdl_enabled() { return pdl_mode() > 0; }

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

// ---- stream-K workspace: one fp32 partial tile + one arrival flag per CTA.  A workspace belongs to ONE stream of work
// at a time (two stream-K GEMMs overlapping on the same slots would consume each other's partials), so it is owned by
// whoever serialises the launches: a UNet handle (its plan runs on one stream at a time), or - for the stand-alone
// entry points - the (device, stream) pair the launch goes to.
bool sk_workspace_alloc(SkWorkspace* w, std::string* err) {
  if (w->ws) return true;
  const int slots = num_sms();
  cudaError_t e = cudaGetDevice(&w->device);
  if (e == cudaSuccess) e = cudaMalloc(&w->ws, (size_t)slots * 128 * 192 * sizeof(float));  // one fp32 tile of the widest kernel per CTA
  if (e == cudaSuccess) e = cudaMalloc(&w->flags, (size_t)slots * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(w->flags, 0, (size_t)slots * sizeof(unsigned));
  if (e != cudaSuccess) {
    if (err) *err = std::string("stream-K workspace: ") + cudaGetErrorString(e);
    sk_workspace_free(w);
    return false;
  }
  w->slots = slots;
  return true;
}
void sk_workspace_free(SkWorkspace* w) {
  if (w->ws) cudaFree(w->ws);
  if (w->flags) cudaFree(w->flags);
  w->ws = nullptr;
  w->flags = nullptr;
  w->slots = 0;
}
const SkWorkspace* sk_workspace_for_stream(cudaStream_t s, std::string* err) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, SkWorkspace> by_stream;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  SkWorkspace& w = by_stream[{dev, s}];
  if (!sk_workspace_alloc(&w, err)) return nullptr;
  return &w;
}
static int sk_min_saving() { return opt(OPT_SK_MIN); }  // k-blocks a launch must save before stream-K pays for its fix-up
int gemm_set_sk_min(int k_blocks) {
  int prev = 0;
  opt_set("sk_min", k_blocks < 0 ? 0 : k_blocks, &prev);
  return prev;
}
static int pair_enabled() { return opt(OPT_GEMM_PAIR); }
int gemm_set_pair(int on) {
  int prev = 0;
  opt_set("gemm_pair", on < 0 ? 0 : on, &prev);
  return prev;
}
// Tile width of a plain GEMM (single row-major K segment; no GEGLU, activation, GroupNorm statistics or forced width).
// 160 where it divides N, else 128 / 64 - and 192 where that saves whole waves of tiles: a wave costs about the bytes a
// CTA pulls per k-block (128 rows of A + bn rows of B), and e.g. the 16x16-latent N = K = 1280 layers of the UNet
// (M = 2560) are 160 tiles = TWO waves on 148 SMs at width 160 but 140 tiles = one wave at width 192.
int gemm_plain_bn(int M, int N) {
  const int base = (N % 160 == 0) ? 160 : (N <= 64 ? 64 : 128);
  if (N < 192) return base;
  const long mt = (M + 127) / 128, sms = num_sms();
  auto cost = [&](int bn) { return ((mt * ((N + bn - 1) / bn) + sms - 1) / sms) * (128 + bn); };
  return cost(192) * 100 < cost(base) * 92 ? 192 : base;
}
int gemm_stats_parts(int N, int M) {  // one part per epilogue column group per N tile of the producer's tile width
  const int bn = gemm_plain_bn(M, N);
  return RCDM_EPI_GROUPS * ((N + bn - 1) / bn);
}

bool gemm_gn_stats_ok(int M, int N, int hw) { return N % 160 == 0 && M % 128 == 0 && hw > 0 && hw % 32 == 0 && M % hw == 0; }

static int pick_bn(const GemmDesc& d) {
  if (d.gn_acc) return 160;
  if (d.force_bn) return d.force_bn;
  if (d.geglu) return geglu_bn(d.N);
  const bool plain = d.nseg == 1 && d.seg[0].mode == SEG_PLAIN && !d.act;
  // (a row-statistics producer must use exactly this width: its consumer sums gemm_stats_parts(N, M) parts; and a
  //  192-wide tile excludes the CTA-pair kernel, whatever the pairing option says)
  if (plain) return gemm_plain_bn(d.M, d.N);
  if (d.N % 160 == 0) return 160;
  if (d.N <= 64) return 64;
  return 128;
}

bool gemm_prepare(const GemmDesc& d, GemmLaunch* l, std::string* err) {
  auto fail = [&](const std::string& m) {
    if (err) *err = "gemm_prepare: " + m;
    return false;
  };
  GemmParams& p = l->p;
  memset(&p, 0, sizeof p);
  memset(&l->maps, 0, sizeof l->maps);
  const int bn = pick_bn(d);
  l->bn = bn;
  l->dt = d.dt;
  p.M = d.M;
  p.N = d.N;
  p.nseg = d.nseg;
  p.out = d.out;
  p.ldo = d.ldo;
  p.bias = d.bias;
  p.res = d.res;
  p.ldr = d.ldr;
  p.geglu = d.geglu;
  p.act = d.act;
  p.stats_out = d.stats_out;
  p.stats_in = d.stats_in;
  p.stats_parts = d.stats_parts;
  p.ln_K = d.Ktot;
  p.ln_c = d.ln_c;
  p.ln_frames = d.ln_frames < 1 ? 1 : d.ln_frames;
  p.ln_rows_per_frame = d.ln_rows_per_frame < 1 ? 1 : d.ln_rows_per_frame;
  p.ln_eps = d.ln_eps;
  if (d.stats_in && (!d.ln_c || d.N % 8 != 0 || (d.geglu && d.N % bn != 0) || d.nseg != 1 ||
                     d.seg[0].mode != SEG_PLAIN))
    return fail("folded LayerNorm needs the c vector, N % 8 == 0 and a single plain K segment");
  if (d.stats_out && (d.geglu || d.N % 8 != 0)) return fail("row statistics need a plain vectorised epilogue");
  p.gn_acc = d.gn_acc;
  p.gn_hw = d.gn_hw;
  l->gn = d.gn_acc ? 1 : 0;
  if (d.gn_acc && (d.geglu || d.act || d.stats_in || !gemm_gn_stats_ok(d.M, d.N, d.gn_hw)))
    return fail("GroupNorm statistics need a plain epilogue, N % 160 == 0, M % 128 == 0 and hw % 32 == 0");
  p.tw = 1;
  p.th = 1;
  p.tn = 128;
  p.tiles_x = 1;
  p.tiles_y = 1;
  if (d.geglu && (d.N % bn != 0)) return fail("GEGLU needs N % BN == 0");
  if (d.Ktot % 8 != 0) return fail("K must be a multiple of 8");
  bool has_conv = false;
  int nmap = 0, kb = 0, kcols = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const ASeg& a = d.seg[s];
    p.seg[s].mode = a.mode;
    p.seg[s].tmap = nmap;
    p.seg[s].cblocks = (a.C + 63) / 64;
    if (d.nseg > 1 && a.C % 64 != 0) return fail("multi-segment K must be 64-aligned");
    if (a.mode == SEG_PLAIN) {
      if (nmap + 1 > 4) return fail("too many tensor maps");
      uint64_t dims[2] = {(uint64_t)a.C, (uint64_t)d.M};
      uint64_t str[1] = {(uint64_t)a.ld * 2};
      uint32_t box[2] = {64, 128};
      if (!encode_tmap(&l->maps.a[nmap], a.ptr, 2, dims, str, box, true, err)) return false;
      nmap += 1;
      kb += p.seg[s].cblocks;
      kcols += a.C;
    } else {
      if (a.C % 64 != 0) return fail("conv channels must be a multiple of 64");
      has_conv = true;
      if (a.mode == SEG_CONV3 || a.mode == SEG_UP2) {
        if (nmap + 1 > 4) return fail("too many tensor maps");
        if (a.H != d.Ho || a.W != d.Wo) return fail("conv3 s1 / folded upsample: input/output grid mismatch");
        if (a.mode == SEG_UP2 && (d.nseg != 1 || d.res)) return fail("folded upsample: single segment, no residual");
        nmap += 1;  // encoded below once the tile geometry is known
      } else {
        if (nmap + 4 > 4) return fail("too many tensor maps");
        if (a.H != 2 * d.Ho || a.W != 2 * d.Wo) return fail("conv3 s2: input must be 2x the output grid");
        nmap += 4;
      }
      kb += seg_taps(a.mode) * p.seg[s].cblocks;
      kcols += seg_taps(a.mode) * a.C;
    }
  }
  if (kcols != d.Ktot) return fail("segment K does not add up to the weight K");
  p.num_kb = kb;
  if (has_conv) {
    if (!is_pow2(d.Ho) || !is_pow2(d.Wo)) return fail("conv path needs power-of-two spatial dims");
    if (d.M != d.NI * d.Ho * d.Wo) return fail("conv M mismatch");
    p.tw = d.Wo < 128 ? d.Wo : 128;
    p.th = d.Ho < 128 / p.tw ? d.Ho : 128 / p.tw;
    p.tn = 128 / (p.tw * p.th);
    p.tiles_x = d.Wo / p.tw;
    p.tiles_y = d.Ho / p.th;
    for (int s = 0; s < d.nseg; ++s) {
      const ASeg& a = d.seg[s];
      const int mi = p.seg[s].tmap;
      uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
      if (a.mode == SEG_CONV3 || a.mode == SEG_UP2) {
        uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.NI};
        uint64_t str[3] = {(uint64_t)a.C * 2, (uint64_t)a.W * a.C * 2, (uint64_t)a.H * a.W * a.C * 2};
        if (!encode_tmap(&l->maps.a[mi], a.ptr, 4, dims, str, box, true, err)) return false;
      } else if (a.mode == SEG_CONV3S2 || a.mode == SEG_CONV3S2A) {
        for (int py = 0; py < 2; ++py)
          for (int px = 0; px < 2; ++px) {
            const char* base = reinterpret_cast<const char*>(a.ptr) + ((size_t)py * a.W + px) * a.C * 2;
            uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)a.W / 2, (uint64_t)a.H / 2, (uint64_t)a.NI};
            uint64_t str[3] = {(uint64_t)2 * a.C * 2, (uint64_t)2 * a.W * a.C * 2, (uint64_t)a.H * a.W * a.C * 2};
            if (!encode_tmap(&l->maps.a[mi + py * 2 + px], base, 4, dims, str, box, true, err)) return false;
          }
      }
    }
  }
  const int m_tiles = has_conv ? ((d.NI + p.tn - 1) / p.tn) * p.tiles_y * p.tiles_x : (d.M + 127) / 128;
  // CTA pairs (cta_group::2, 256-row tiles: 28 % fewer operand bytes per SM).  Measured on B200 with the two-producer
  // kernel: K = 320 layers lose 10-15 %, plain K = 640 / 1280 layers are unchanged, GEGLU projections with K >= 640 gain
  // 6-7 %, 3x3 convolutions gain 8-10 % from K = 2880 on.
  // OPT_GEMM_PAIR = 0 never, 1 heuristic (default), 2 whenever there are >= 2 M tiles.
  const int pair_mode = pair_enabled();
  l->pair = (!d.no_pair && !d.act && bn != 192 && num_sms() >= 2 &&  // activation epilogues / 192-wide tiles: single-CTA kernel only
             ((d.force_pair && m_tiles >= 2) || (pair_mode == 2 && m_tiles >= 2) ||
              (pair_mode == 1 && m_tiles >= 16 && (kb >= 90 || (has_conv && kb >= 40) || (d.geglu && kb >= 10))))) ? 1 : 0;
  {
    uint64_t dims[2] = {(uint64_t)d.Ktot, (uint64_t)d.w_rows};
    uint64_t str[1] = {(uint64_t)d.Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)(l->pair ? bn / 2 : bn)};
    if (!encode_tmap(&l->maps.b, d.w, 2, dims, str, box, true, err)) return false;
  }
  // ---- epilogue through the staging buffers + TMA: needs 16-byte aligned rows of the output / residual
  {
    const int n_out = d.geglu ? d.N / 2 : d.N;
    const int wcols = (d.geglu ? bn / 2 : bn) / RCDM_EPI_GROUPS;
    const bool ok = (d.N % 8 == 0) && (d.ldo % 8 == 0) && (!d.res || d.ldr % 8 == 0) && n_out % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 &&
                    (!d.res || (reinterpret_cast<uintptr_t>(d.res) & 15) == 0) && (wcols * 2) % 16 == 0;
    p.epi_tma = ok ? 1 : 0;
    const bool up2 = d.seg[0].mode == SEG_UP2;
    p.up_py = d.up_py;
    p.up_px = d.up_px;
    p.out4d = up2 ? 1 : 0;
    if (up2 && !ok) return fail("folded upsample needs the TMA epilogue (N % 8 == 0, aligned rows)");
    if (ok && up2) {
      // class (py, px) of the [NI, 2 Ho, 2 Wo, ldo] output: pixel (2 y + py, 2 x + px) <-> class-grid pixel (y, x)
      const char* base = reinterpret_cast<const char*>(d.out) + ((size_t)d.up_py * 2 * d.Wo + d.up_px) * d.ldo * 2;
      uint64_t dims[4] = {(uint64_t)n_out, (uint64_t)d.Wo, (uint64_t)d.Ho, (uint64_t)d.NI};
      uint64_t str[3] = {(uint64_t)2 * d.ldo * 2, (uint64_t)2 * (2 * d.Wo) * d.ldo * 2, (uint64_t)(2 * d.Ho) * (2 * d.Wo) * d.ldo * 2};
      uint32_t box[4] = {(uint32_t)wcols, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
      if (!encode_tmap(&l->maps.o, base, 4, dims, str, box, false, err)) return false;
    } else if (ok) {
      uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)d.M};
      uint32_t box[2] = {(uint32_t)wcols, 128};
      uint64_t str[1] = {(uint64_t)d.ldo * 2};
      if (!encode_tmap(&l->maps.o, d.out, 2, dims, str, box, false, err)) return false;
      if (d.res) {
        uint64_t strr[1] = {(uint64_t)d.ldr * 2};
        if (!encode_tmap(&l->maps.r, d.res, 2, dims, strr, box, false, err)) return false;
      }
    }
    if (!ok && (d.stats_out || d.stats_in || d.geglu || d.act || d.gn_acc))
      return fail("this epilogue needs 16-byte aligned output rows");
    if (d.act && d.geglu) return fail("act (plain GELU) and geglu are mutually exclusive");
  }
  p.num_m_tiles = l->pair ? (m_tiles + 1) / 2 : m_tiles;  // pair mode: counted in 256-row tile pairs
  p.num_n_tiles = (d.N + bn - 1) / bn;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int sms = l->pair ? num_sms() / 2 : num_sms();    // workers: CTAs or CTA pairs
  const int cta_per_worker = l->pair ? 2 : 1;
  l->grid = dim3((tiles < sms ? tiles : sms) * cta_per_worker, 1, 1);
  // ---- stream-K when data-parallel tiling leaves the last wave (or most of the GPU) idle
  p.sk = 0;
  const bool vec_ok = p.epi_tma != 0;
  const int min_saving = sk_min_saving();
  if (min_saving > 0 && vec_ok && !d.no_sk && d.sk && d.sk->ws && sms * cta_per_worker <= d.sk->slots && tiles % sms != 0) {
    const double waves = (double)tiles / sms;
    const double saving_kb = (std::ceil(waves) - waves) * p.num_kb;
    if (saving_kb >= min_saving && (long long)tiles * p.num_kb >= 4LL * sms) {
      p.sk = 1;
      p.sk_ws = d.sk->ws;
      p.sk_flags = d.sk->flags;
      l->grid = dim3(sms * cta_per_worker, 1, 1);
    }
  }
  return true;
}

template <typename T, int BN> static void launch_one(const GemmLaunch& l, cudaStream_t s) {
  if constexpr (BN != 192) {
    if (l.p.act) {  // GELU / SiLU epilogue (stage-1 prior): separate instantiation, see gemm_tcgen05.cuh
      launch_k(gemm_tcgen05_kernel<T, BN, false, true>, l.grid, dim3(GemmCfg<BN, false>::THREADS),
               GemmCfg<BN, false>::SMEM_BYTES, s, l.maps, l.p);
      return;
    }
  }
  if (!l.pair && l.p.sk && !l.gn) {
    // stream-K: CTAs wait for each other's partial tiles through flags in global memory -> cooperative launch
    launch_coop(gemm_tcgen05_kernel<T, BN, false>, l.grid, dim3(GemmCfg<BN, false>::THREADS), GemmCfg<BN, false>::SMEM_BYTES, s,
                l.maps, l.p);
    return;
  }
  if (!l.pair) {
    if constexpr (BN == 160) {
      if (l.gn) {
        launch_k(gemm_tcgen05_kernel<T, BN, false, false, true>, l.grid, dim3(GemmCfg<BN, false>::THREADS),
                 GemmCfg<BN, false>::SMEM_BYTES, s, l.maps, l.p);
        return;
      }
    }
    launch_k(gemm_tcgen05_kernel<T, BN, false>, l.grid, dim3(GemmCfg<BN, false>::THREADS), GemmCfg<BN, false>::SMEM_BYTES, s,
             l.maps, l.p);
    return;
  }
  if constexpr (BN != 192) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = l.grid;
  cfg.blockDim = dim3(GemmCfg<BN, true>::THREADS);
  cfg.dynamicSmemBytes = GemmCfg<BN, true>::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if constexpr (BN == 160) {
    if (l.gn) {
      cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<T, BN, true, false, true>, l.maps, l.p);
      return;
    }
  }
  cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<T, BN, true>, l.maps, l.p);
  }
}

void gemm_launch(const GemmLaunch& l, cudaStream_t s) {
  if (l.dt == DT_F16) {
    if (l.bn == 64) launch_one<__half, 64>(l, s);
    else if (l.bn == 128) launch_one<__half, 128>(l, s);
    else if (l.bn == 192) launch_one<__half, 192>(l, s);
    else launch_one<__half, 160>(l, s);
  } else {
    if (l.bn == 64) launch_one<__nv_bfloat16, 64>(l, s);
    else if (l.bn == 128) launch_one<__nv_bfloat16, 128>(l, s);
    else if (l.bn == 192) launch_one<__nv_bfloat16, 192>(l, s);
    else launch_one<__nv_bfloat16, 160>(l, s);
  }
}

template <typename T, int BN> static cudaError_t set_attr() {
  cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<T, BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GemmCfg<BN, false>::SMEM_BYTES);
  if constexpr (BN != 192) {  // (192-wide tiles: single-CTA plain kernel only)
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tcgen05_kernel<T, BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GemmCfg<BN, true>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tcgen05_kernel<T, BN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GemmCfg<BN, false>::SMEM_BYTES);
  }
  if constexpr (BN == 160) {
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tcgen05_kernel<T, BN, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GemmCfg<BN, false>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tcgen05_kernel<T, BN, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GemmCfg<BN, true>::SMEM_BYTES);
  }
  return e;
}

bool gemm_setup_attributes(std::string* err) {
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = set_attr<__half, 64>();
  if (e == cudaSuccess) e = set_attr<__half, 128>();
  if (e == cudaSuccess) e = set_attr<__half, 160>();
  if (e == cudaSuccess) e = set_attr<__half, 192>();
  if (e == cudaSuccess) e = set_attr<__nv_bfloat16, 192>();
  if (e == cudaSuccess) e = set_attr<__nv_bfloat16, 64>();
  if (e == cudaSuccess) e = set_attr<__nv_bfloat16, 128>();
  if (e == cudaSuccess) e = set_attr<__nv_bfloat16, 160>();
  if (e != cudaSuccess) {
    if (err) *err = std::string("cudaFuncSetAttribute(gemm): ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// CUDA-core debug kernel: one thread per output element, same segment / epilogue semantics
// ------------------------------------------------------------------------------------------
struct SimpleArgs {
  GemmDesc d;
  int bn;
};

template <typename T>
__device__ float simple_dot(const GemmDesc& d, int m, int wrow) {
  const T* w = reinterpret_cast<const T*>(d.w) + (size_t)wrow * d.Ktot;
  float acc = 0.f;
  int koff = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const ASeg& a = d.seg[s];
    const T* ap = reinterpret_cast<const T*>(a.ptr);
    if (a.mode == SEG_PLAIN) {
      for (int k = 0; k < a.C; ++k) acc += DT<T>::to_f(ap[(size_t)m * a.ld + k]) * DT<T>::to_f(w[koff + k]);
      koff += a.C;
    } else {
      const int x = m % d.Wo, y = (m / d.Wo) % d.Ho, n = m / (d.Wo * d.Ho);
      const int st = (a.mode == SEG_CONV3S2 || a.mode == SEG_CONV3S2A) ? 2 : 1;
      const int off = a.mode == SEG_CONV3S2A ? 0 : 1;  // padding (0,1,0,1): taps start at the pixel itself
      for (int tap = 0; tap < 9; ++tap) {
        const int sy = y * st + tap / 3 - off, sx = x * st + tap % 3 - off;
        if (sy >= 0 && sy < a.H && sx >= 0 && sx < a.W) {
          const T* src = ap + (((size_t)n * a.H + sy) * a.W + sx) * a.C;
          for (int c = 0; c < a.C; ++c) acc += DT<T>::to_f(src[c]) * DT<T>::to_f(w[koff + tap * a.C + c]);
        }
      }
      koff += 9 * a.C;
    }
  }
  return acc;
}

template <typename T>
__global__ void gemm_simple_kernel(const SimpleArgs sa) {
  const GemmDesc& d = sa.d;
  const int nout = d.geglu ? d.N / 2 : d.N;
  const size_t total = (size_t)d.M * nout;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx / nout), j = (int)(idx % nout);
    float v;
    if (d.geglu) {
      const int hb = sa.bn / 2;
      const int rh = (j / hb) * sa.bn + (j % hb), rg = rh + hb;
      float hv = simple_dot<T>(d, m, rh), gv = simple_dot<T>(d, m, rg);
      if (d.bias) {
        hv += d.bias[rh];
        gv += d.bias[rg];
      }
      v = hv * gelu_erf_f(gv);
    } else {
      v = simple_dot<T>(d, m, j);
      if (d.bias) v += d.bias[j];
      if (d.act == 1) v = gelu_erf_f(v);
      else if (d.act == 2) v = silu_f(v);
      if (d.res) v += DT<T>::to_f(reinterpret_cast<const T*>(d.res)[(size_t)m * d.ldr + j]);
    }
    reinterpret_cast<T*>(d.out)[(size_t)m * d.ldo + j] = DT<T>::from_f(v);
  }
}

void gemm_simple_launch(const GemmDesc& d, cudaStream_t s) {
  SimpleArgs sa;
  sa.d = d;
  sa.bn = pick_bn(d);
  const size_t total = (size_t)d.M * (d.geglu ? d.N / 2 : d.N);
  const int blocks = (int)((total + 255) / 256 < 65535 * 16 ? (total + 255) / 256 : 65535 * 16);
  if (d.dt == DT_F16) gemm_simple_kernel<__half><<<blocks, 256, 0, s>>>(sa);
  else gemm_simple_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(sa);
}

// ------------------------------------------------------------------------------------------
// fused GEGLU feed-forward (ffn_fused.cuh)
// ------------------------------------------------------------------------------------------
bool ffn_prepare(const FfnDesc& d, FfnLaunch* l, std::string* err) {
  auto fail = [&](const std::string& m) {
    if (err) *err = "ffn_prepare: " + m;
    return false;
  };
  if (d.dt != DT_F16 && d.dt != DT_BF16) return fail("dtype must be f16/bf16");
  if (!d.y || !d.stats_in || !d.w1f || !d.c1 || !d.w2 || !d.out || d.M <= 0 || d.stats_parts <= 0) return fail("bad argument");
  constexpr int C = FfnCfg::C, J = FfnCfg::J;
  memset(&l->maps, 0, sizeof l->maps);
  memset(&l->p, 0, sizeof l->p);
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)d.M};
    uint64_t str[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {64, 128};
    if (!encode_tmap_sw(&l->maps.a1, d.y, 2, dims, str, box, 128, err)) return false;
  }
  const int pair = d.pair && num_sms() >= 2 ? 1 : 0;
  {  // W1f [2J, C] seen as (64 columns, 2J rows, C / 64 k-blocks): one box = a chunk's 64 rows (pair: this CTA's 32) x all
     // k-blocks
    uint64_t dims[3] = {64, (uint64_t)2 * J, (uint64_t)C / 64};
    uint64_t str[2] = {(uint64_t)C * 2, 128};
    uint32_t box[3] = {64, (uint32_t)(pair ? 32 : 64), (uint32_t)C / 64};
    if (!encode_tmap_sw(&l->maps.b1, d.w1f, 3, dims, str, box, 128, err)) return false;
  }
  {
    uint64_t dims[2] = {(uint64_t)J, (uint64_t)C};
    uint64_t str[1] = {(uint64_t)J * 2};
    uint32_t box[2] = {(uint32_t)FfnCfg::CH, (uint32_t)(pair ? 80 : 160)};
    if (!encode_tmap_sw(&l->maps.b2, d.w2, 2, dims, str, box, 64, err)) return false;
  }
  FfnParams& p = l->p;
  p.M = d.M;
  p.num_m_tiles = (d.M + 127) / 128;
  p.stats_in = d.stats_in;
  p.stats_parts = d.stats_parts;
  p.ln_eps = d.ln_eps;
  p.c1 = d.c1;
  p.bias2 = d.bias2;
  p.res = d.y;
  p.out = d.out;
  l->dt = d.dt;
  l->pair = pair;
  const int sms = num_sms();
  if (pair) {
    const int pairs = (p.num_m_tiles + 1) / 2, workers = sms / 2;
    l->grid = dim3(2 * (pairs < workers ? pairs : workers));
  } else {
    l->grid = dim3(p.num_m_tiles < sms ? p.num_m_tiles : sms);
  }
  return true;
}
void ffn_launch(const FfnLaunch& l, cudaStream_t s) {
  if (l.pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = l.grid;
    cfg.blockDim = dim3(FfnPairCfg::THREADS);
    cfg.dynamicSmemBytes = FfnPairCfg::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (l.dt == DT_F16) cudaLaunchKernelEx(&cfg, ffn_geglu_fused_pair_kernel<__half>, l.maps, l.p);
    else cudaLaunchKernelEx(&cfg, ffn_geglu_fused_pair_kernel<__nv_bfloat16>, l.maps, l.p);
    return;
  }
  if (l.dt == DT_F16)
    launch_k(ffn_geglu_fused_kernel<__half>, l.grid, dim3(FfnCfg::THREADS), FfnCfg::SMEM_BYTES, s, l.maps, l.p);
  else
    launch_k(ffn_geglu_fused_kernel<__nv_bfloat16>, l.grid, dim3(FfnCfg::THREADS), FfnCfg::SMEM_BYTES, s, l.maps, l.p);
}
bool ffn_setup_attributes(std::string* err) {
  cudaError_t e = cudaFuncSetAttribute(ffn_geglu_fused_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       FfnCfg::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(ffn_geglu_fused_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             FfnCfg::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(ffn_geglu_fused_pair_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             FfnPairCfg::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(ffn_geglu_fused_pair_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             FfnPairCfg::SMEM_BYTES);
  if (e != cudaSuccess) {
    if (err) *err = std::string("cudaFuncSetAttribute(ffn): ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

}  // namespace rcdm

#if RCDM_GEMM_TRACE
// variant builds only: copy the GEMM timeline stamps to the host
extern "C" __attribute__((visibility("default"))) int rcdm_debug_gemm_trace_read(long long* stamps) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(stamps, rcdm::g_gemm_trace, sizeof(rcdm::g_gemm_trace)) == cudaSuccess ? 0 : 1;
}
#endif

#if RCDM_FFN_TRACE
extern "C" __attribute__((visibility("default"))) int rcdm_debug_ffn_trace_read(long long* stamps) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(stamps, rcdm::g_ffn_trace, sizeof(rcdm::g_ffn_trace)) == cudaSuccess ? 0 : 1;
}
#endif
