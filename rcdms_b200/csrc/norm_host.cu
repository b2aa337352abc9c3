// Host launchers for the normalisation kernels (shared by the UNet plan and the single-kernel C entry points).
#include "internal.h"

namespace rcdm {

size_t gn_scratch_bytes(int nstat, int groups) {
  // partial [nstat][<=1024 chunks][groups] float2 is bounded by chunks*nstat <= 1024 + nstat
  return 65536 + (size_t)(1024 + 2 * (size_t)nstat) * groups * 8 + (size_t)nstat * groups * 8 + 2048;
}

void gn_configure(GnLaunch* l, int dt, const void* x0, int C0, const void* x1, int C1, int rows, int rows_per_stat,
                  int groups, float eps, const float* gamma, const float* beta, void* out, int silu, void* scratch) {
  memset(l, 0, sizeof *l);
  GnArgs& a = l->a;
  const int C = C0 + C1;
  a.x0 = x0;
  a.x1 = x1;
  a.C0 = C0;
  a.C1 = C1;
  a.groups = groups;
  a.rows_per_stat = rows_per_stat;
  a.nstat = rows / rows_per_stat;
  a.eps = eps;
  a.gamma = gamma;
  a.beta = beta;
  a.out = out;
  a.silu = silu;
  int target = 592 / a.nstat;
  if (target < 1) target = 1;
  if (target > 1024) target = 1024;
  int chunks = 1;
  for (int c = target; c >= 1; --c)
    if (rows_per_stat % c == 0) {
      chunks = c;
      break;
    }
  a.rows_per_cta = rows_per_stat / chunks;
  // fixed layout: [counters | stats | partial] so the (always-zero-at-rest) counters never alias partial sums
  unsigned char* s = reinterpret_cast<unsigned char*>(scratch);
  a.counters = reinterpret_cast<unsigned*>(s);
  size_t off = 65536;  // counters region is fixed-size (<= 16384 statistic batches) for every caller
  a.stats = reinterpret_cast<float2*>(s + off);
  off += ((size_t)a.nstat * groups * 8 + 1023) & ~size_t(1023);
  a.partial = reinterpret_cast<float2*>(s + off);
  const int vecs = C / 8;
  int k = 256 / vecs;
  if (k > a.rows_per_cta) k = a.rows_per_cta;
  if (k < 1) k = 1;
  while (vecs * k < groups) ++k;  // the group reduction needs >= groups threads
  l->threads = vecs * k;
  l->smem = (size_t)k * 2 * C * 4;
  l->grid = dim3(chunks, a.nstat);
  l->total_vecs = (size_t)rows * vecs;
  size_t g = (l->total_vecs + 255) / 256;
  l->agrid = (int)(g < 148 * 8 ? (g ? g : 1) : 148 * 8);
  l->dt = dt;
  l->rows_per_cta_2k = a.rows_per_cta;
  // ---- fused single-launch path
  const bool fused_on = opt(OPT_GN_FUSED) != 0;
  l->fused = 0;
  if (fused_on && vecs <= 512 && a.nstat <= 8192 && groups <= 64) {
    const int sms = num_sms();
    int kf = 512 / vecs;
    if (kf < 1) kf = 1;
    int cps = sms / a.nstat;
    if (cps < 1) cps = 1;
    if (cps > rows_per_stat) cps = rows_per_stat;
    const int rpc = (rows_per_stat + cps - 1) / cps;
    cps = (rows_per_stat + rpc - 1) / rpc;  // drop CTAs that would own no rows
    if (kf > rpc) kf = rpc;
    size_t scratch_b = (size_t)kf * 2 * C * 4;
    if (scratch_b < (size_t)cps * groups * 8) scratch_b = (size_t)cps * groups * 8;  // also holds the batch's partials
    scratch_b = (scratch_b + 15) & ~size_t(15);
    const size_t budget = 232448 - 1024 - scratch_b;
    size_t cache_rows = budget / ((size_t)C * 2);
    if (cache_rows > (size_t)rpc) cache_rows = rpc;
    l->fused = 1;
    l->cps = cps;
    l->cache_rows = (int)cache_rows;
    l->fthreads = (vecs * kf + 31) / 32 * 32;
    l->fk = kf;
    l->fgrid = a.nstat * cps;
    l->fsmem = scratch_b + cache_rows * (size_t)C * 2;
    a.rows_per_cta = rpc;
  }
}

void gn_configure_from_stats(GnLaunch* l, int dt, const void* x0, int C0, const unsigned long long* acc0, const void* x1,
                             int C1, const unsigned long long* acc1, int rows, int rows_per_stat, int hw, int groups,
                             float eps, const float* gamma, const float* beta, void* out, int silu) {
  memset(l, 0, sizeof *l);
  GnArgs& a = l->a;
  const int C = C0 + C1;
  a.x0 = x0;
  a.x1 = x1;
  a.C0 = C0;
  a.C1 = C1;
  a.groups = groups;
  a.rows_per_stat = rows_per_stat;
  a.nstat = rows / rows_per_stat;
  a.eps = eps;
  a.gamma = gamma;
  a.beta = beta;
  a.out = out;
  a.silu = silu;
  a.acc0 = acc0;
  a.acc1 = acc1;
  a.hw = hw;
  // pure streaming pass: 2 CTAs per SM (what the kernel's registers allow), every CTA inside one statistic batch; the
  // chunk count need not divide the batch (the kernel clips the last chunk)
  int chunks = (num_sms() * 2) / a.nstat;
  if (chunks < 1) chunks = 1;
  if (chunks > rows_per_stat) chunks = rows_per_stat;
  a.rows_per_cta = (rows_per_stat + chunks - 1) / chunks;
  chunks = (rows_per_stat + a.rows_per_cta - 1) / a.rows_per_cta;  // drop CTAs that would own no rows
  const int vecs = C / 8;
  // k row lanes: ~320 threads, at most 384, at least the 8 * groups threads of the statistics prologue
  int k = 320 / vecs;
  if (k < 1) k = 1;
  while (vecs * (k + 1) <= 384 && vecs * k < 288) ++k;
  l->threads = (vecs * k + 31) / 32 * 32;
  if (l->threads < 8 * groups) l->threads = 8 * groups;
  l->grid = dim3(chunks, a.nstat);
  l->dt = dt;
  l->from_stats = 1;
}

bool gn_setup_attributes(std::string* err) {
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gn_fused_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gn_fused_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024);
  if (e != cudaSuccess) {
    if (err) *err = std::string("cudaFuncSetAttribute(gn_fused): ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

void gn_run(const GnLaunch& l, cudaStream_t s) {
  if (l.from_stats) {
    if (l.dt == DT_F16) launch_k(gn_apply_stats_kernel<__half>, l.grid, dim3(l.threads), 0, s, l.a);
    else launch_k(gn_apply_stats_kernel<__nv_bfloat16>, l.grid, dim3(l.threads), 0, s, l.a);
    g_launches += 1;
    return;
  }
  if (l.fused) {
    // cps > 1: the CTAs of a statistic batch meet at a barrier in global memory -> cooperative launch (co-residency
    // guaranteed by the driver); cps == 1: no inter-CTA dependency, ordinary launch
    if (l.cps > 1) {
      if (l.dt == DT_F16)
        launch_coop(gn_fused_kernel<__half>, dim3(l.fgrid), dim3(l.fthreads), l.fsmem, s, l.a, l.cps, l.cache_rows, l.fk);
      else
        launch_coop(gn_fused_kernel<__nv_bfloat16>, dim3(l.fgrid), dim3(l.fthreads), l.fsmem, s, l.a, l.cps, l.cache_rows, l.fk);
    } else if (l.dt == DT_F16)
      launch_k(gn_fused_kernel<__half>, dim3(l.fgrid), dim3(l.fthreads), l.fsmem, s, l.a, l.cps, l.cache_rows, l.fk);
    else
      launch_k(gn_fused_kernel<__nv_bfloat16>, dim3(l.fgrid), dim3(l.fthreads), l.fsmem, s, l.a, l.cps, l.cache_rows, l.fk);
    g_launches += 1;
    return;
  }
  if (l.dt == DT_F16) {
    launch_k(gn_stats_kernel<__half>, l.grid, dim3(l.threads), l.smem, s, l.a);
    launch_k(gn_apply_kernel<__half>, l.grid, dim3(l.threads), 0, s, l.a);
  } else {
    launch_k(gn_stats_kernel<__nv_bfloat16>, l.grid, dim3(l.threads), l.smem, s, l.a);
    launch_k(gn_apply_kernel<__nv_bfloat16>, l.grid, dim3(l.threads), 0, s, l.a);
  }
  g_launches += 2;
}

bool ln_run(int dt, const void* x, void* o, const float* gp, const float* bp, int nrows, int C, float eps,
            const float* pep, int rows_per_frame, int frames, cudaStream_t s) {
  const int maxv = (C / 8 + 31) / 32;
  if (maxv > 8 || C % 8) return false;  // C <= 2048 (the stage-1 prior's width)
  if (maxv > 5) {  // wide rows (the stage-1 prior, C = 2048): one CTA per row
    const bool wide_on = opt(OPT_LN_WIDE) != 0;
    if (wide_on) {
      if (dt == DT_F16)
        launch_k(layernorm_wide_kernel<__half>, dim3(nrows), dim3(128), 0, s, reinterpret_cast<const __half*>(x),
                 reinterpret_cast<__half*>(o), gp, bp, nrows, C, eps, pep, rows_per_frame, frames);
      else
        launch_k(layernorm_wide_kernel<__nv_bfloat16>, dim3(nrows), dim3(128), 0, s,
                 reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(o), gp, bp, nrows, C, eps,
                 pep, rows_per_frame, frames);
      g_launches++;
      return true;
    }
  }
  // rows per warp: keep ~8 16-byte loads in flight per lane
  const int rpw = maxv <= 1 ? 8 : maxv == 2 ? 4 : maxv <= 5 ? 2 : 1;
  const int warps = (nrows + rpw - 1) / rpw;
  const int blocks = (warps + 7) / 8;
#define RCDM_LN(T, V, R)                                                                                           \
  launch_k(layernorm_kernel<T, V, R>, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const T*>(x),                \
           reinterpret_cast<T*>(o), gp, bp, nrows, C, eps, pep, rows_per_frame, frames)
  if (dt == DT_F16) {
    if (maxv <= 1) RCDM_LN(__half, 1, 8);
    else if (maxv == 2) RCDM_LN(__half, 2, 4);
    else if (maxv == 3) RCDM_LN(__half, 3, 2);
    else if (maxv <= 5) RCDM_LN(__half, 5, 2);
    else RCDM_LN(__half, 8, 1);
  } else {
    if (maxv <= 1) RCDM_LN(__nv_bfloat16, 1, 8);
    else if (maxv == 2) RCDM_LN(__nv_bfloat16, 2, 4);
    else if (maxv == 3) RCDM_LN(__nv_bfloat16, 3, 2);
    else if (maxv <= 5) RCDM_LN(__nv_bfloat16, 5, 2);
    else RCDM_LN(__nv_bfloat16, 8, 1);
  }
#undef RCDM_LN
  g_launches++;
  return true;
}

}  // namespace rcdm
