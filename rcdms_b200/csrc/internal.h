// Internal (non-ABI) declarations shared by the translation units of librcdm_b200.so.
#pragma once
#include <atomic>
#include <string>

#include "launch.h"
#include "model.h"
#include "norm_kernels.cuh"

namespace rcdm {
extern thread_local std::string g_err;
extern std::atomic<uint64_t> g_launches;
int set_err(const std::string& m);

int unet_create(const rcdm_unet_config* cfg, rcdm_unet** out);
void unet_destroy(rcdm_unet* h);
int unet_load_weight(rcdm_unet* h, const char* name, const void* data, int dtype, const int64_t* dims, int ndim,
                     void* stream);
int unet_load_weights(rcdm_unet* h, int count, const char* const* names, const void* const* data, const int* dtypes,
                      const int64_t* dims, const int* ndims, void* stream);
int unet_prepare(rcdm_unet* h, int batch, int frames, int height, int width, int ctx_len);
int unet_run(rcdm_unet* h, bool run_ctx, bool run_step, cudaStream_t st);
// wf [C, 5C] = [wp | wp w2] (16 bit, fp32 accumulation), cf [C] = wp b2 + bp: proj_out folded over ff.net.2 (unet.cu)
void fold_proj_launch(int dt, const void* wp, const void* w2, const float* b2, const float* bp, void* wf, float* cf, int C,
                      cudaStream_t st);

struct GnLaunch {
  GnArgs a;
  dim3 grid;
  int threads;
  size_t smem;
  size_t total_vecs;
  int agrid, dt;
  // fused single-launch path (gn_fused_kernel); fused == 0 -> the two-kernel path above
  int fused, cps, cache_rows, fthreads, fgrid, fk;
  size_t fsmem;
  int rows_per_cta_2k;  // rows_per_cta of the two-kernel path (a.rows_per_cta is the fused value when fused)
  int from_stats;       // 1: statistics come from the producing GEMMs' epilogues -> gn_apply_stats_kernel alone
};
int num_sms();
bool gn_setup_attributes(std::string* err);
size_t gn_scratch_bytes(int nstat, int groups);
// scratch must be zero-initialised once (the kernels leave the counters zero)
void gn_configure(GnLaunch* l, int dt, const void* x0, int C0, const void* x1, int C1, int rows, int rows_per_stat,
                  int groups, float eps, const float* gamma, const float* beta, void* out, int silu, void* scratch);
// statistics from the producers' epilogue accumulators (acc1 / C1 = 0: single tensor); hw = rows per image
void gn_configure_from_stats(GnLaunch* l, int dt, const void* x0, int C0, const unsigned long long* acc0, const void* x1,
                             int C1, const unsigned long long* acc1, int rows, int rows_per_stat, int hw, int groups,
                             float eps, const float* gamma, const float* beta, void* out, int silu);
void gn_run(const GnLaunch& l, cudaStream_t s);
bool ln_run(int dt, const void* x, void* out, const float* gamma, const float* beta, int rows, int C, float eps,
            const float* pe, int rows_per_frame, int frames, cudaStream_t s);
}  // namespace rcdm
