// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M, N] = sum over K-segments of A_seg[M, K_seg] * W[N, K_total]^T   (+ epilogue)
//
// Persistent kernel, one CTA per SM, looping over 128 x BN output tiles.  Warp roles:
//   warp 0, 11 TMA producers (one elected lane each; they take the k-blocks alternately, see RCDM_GEMM_PRODUCERS)
//                                                : HBM/L2 -> 128B-swizzled smem ring (4-6 stages), run ahead
//                                                  across tile boundaries
//   warp 1     MMA issuer    (one elected lane)  : tcgen05.mma into one of TWO TMEM accumulators; owns TMEM alloc
//   warps 2-9  epilogue (8 warps)                : tcgen05.ld -> scale * acc + vector (+GEGLU) (+residual tile, which the
//                                                  producer TMA-loaded into the staging buffer a tile ahead)
//                                                  -> 16-bit staging buffer (2 buffers)
//   warp 10    store warp    (one elected lane)  : waits for a staging buffer to be complete, issues ONE TMA store per column
//                                                  group, waits until the store has read the buffer, then TMA-loads the
//                                                  residual tile that will use this buffer next (two tiles ahead)
// so the main loop of tile i+1 overlaps the epilogue of tile i, and no epilogue thread ever waits for a store to drain
// (measured on the K = 320 GEMMs: with the store issued and drained by an epilogue thread, the per-tile chain
//  "store drained -> this thread's math -> barrier -> store" cost 20 % of the launch).
//
// The K loop runs over up to three "segments", each reading A through its own TMA tensor map(s):
//   SEG_PLAIN    A is a row-major [M, K] matrix             (Linear, conv1x1, im2col'ed conv_in)
//   SEG_CONV3    A is an NHWC activation; 3x3, stride 1, pad 1: for each tap the SAME 4-D tensor map is
//                read at spatial offset (kx-1, ky-1); TMA out-of-bounds zero fill implements the padding,
//                so no im2col buffer ever exists ("im2col-free" implicit GEMM)
//   SEG_CONV3S2  3x3, stride 2, pad 1: four parity-subsampled 4-D maps (y%2, x%2), tap -> (map, offset)
//   SEG_UP2      nearest-2x upsample FOLDED into the following 3x3 / pad 1 conv (Upsample3D, resnet.py:32-80): the output
//                pixels of parity (py, px) see only 2 x 2 distinct input pixels, (y + ty - 1 + py, x + tx - 1 + px), with
//                the 3x3 weights that fall on the same input pixel summed beforehand -> one launch per parity class on
//                the ORIGINAL activation (K = 4 Cin instead of 9 Cin, no 4x tensor), output through a strided 4-D map
//   SEG_CONV3S2A 3x3, stride 2, padding (0,1,0,1) (right / bottom only: diffusers Downsample2D with padding=0, the VAE
//                encoder): same four maps, tap row = 2y + ky -> parity ky & 1, offset ky >> 1
// Segments accumulate into the same TMEM tile, which is how conv2 + the 1x1 shortcut of a resblock (whose input
// is the un-materialised concat [h | skip]) become a single launch.
#pragma once
#include <type_traits>
#include "common.cuh"

// Compile-time experiment switch for bottleneck analysis (never set in a product build; scripts/build_variants.sh):
//   3 = producer loads the A tile only for the first k-block of a tile (B still streams), 4 = neither A nor B after
//   the first k-block of a tile (MMAs run on stale smem): isolates epilogue + MMA issue.
//   5 = the epilogue only waits for the accumulator and releases it (no TMEM loads, math, residual or stores),
//   6 = full epilogue math but no residual TMA loads and no TMA stores, 7 = B is loaded only for the CTA's first tile
//   ("weight-stationary" emulation), 9 = 4 + 5 (MMA issue alone).
// (Measured with the earlier per-thread-store epilogue: the slab -> registers -> global "phase 2" cost 27 % of a
//  K = 320 GEMM and the TMEM loads nothing, which is why the epilogue now ends in TMA stores.)
#ifndef RCDM_GEMM_EXPERIMENT
#define RCDM_GEMM_EXPERIMENT 0
#endif
// Column groups of the epilogue: 4 * RCDM_EPI_GROUPS epilogue warps, each owning 32 rows x BN / RCDM_EPI_GROUPS
// accumulator columns.  Measured on B200: 4 groups (16 warps, 96 registers per thread, spills) are 10-35 % SLOWER
// than 2 groups (8 warps, 168 registers) on the small-K GEMMs, so 2 is the product setting.
#ifndef RCDM_EPI_GROUPS
#define RCDM_EPI_GROUPS 2
#endif
// TMA producer threads (one elected lane each of RCDM_GEMM_PRODUCERS warps; k-block g of a CTA is loaded by producer
// g % RCDM_GEMM_PRODUCERS).  Measured with scripts/micro/tma_box_bench.cu on the K = 320 access pattern (A 16 KB + B 20 KB
// boxes, L2-resident): ONE issuing thread sustains 34 B/clk/SM whatever the ring depth (each cp.async.bulk.tensor costs its
// issuing thread ~500 cycles before the next one leaves), TWO threads in different warps 58-60 B/clk/SM - and the 128 x 160
// tile needs 36 KB per 320 MMA cycles.
#ifndef RCDM_GEMM_PRODUCERS
#define RCDM_GEMM_PRODUCERS 2
#endif

// Timeline instrumentation (variant builds only, -DRCDM_GEMM_TRACE=1; scripts/gemm_trace.py): clock64 stamps of the hand-offs
// between the warp roles for the first 16 tiles of every CTA, read back through rcdm_debug_gemm_trace_read.
#ifndef RCDM_GEMM_TRACE
#define RCDM_GEMM_TRACE 0
#endif

namespace rcdm {

#if RCDM_GEMM_TRACE
constexpr int GEMM_TRACE_CTAS = 148, GEMM_TRACE_TILES = 16;
__device__ long long g_gemm_trace[GEMM_TRACE_CTAS * GEMM_TRACE_TILES * 16];
#define GEMM_STAMP(it, slot)                                                                                   \
  do {                                                                                                         \
    if (blockIdx.x < GEMM_TRACE_CTAS && (it) < GEMM_TRACE_TILES)                                               \
      g_gemm_trace[((size_t)blockIdx.x * GEMM_TRACE_TILES + (it)) * 16 + (slot)] = clock64();                  \
  } while (0)
#else
#define GEMM_STAMP(it, slot) do { } while (0)
#endif

enum : int { SEG_PLAIN = 0, SEG_CONV3 = 1, SEG_CONV3S2 = 2, SEG_CONV3S2A = 3, SEG_UP2 = 4 };
__host__ __device__ constexpr int seg_taps(int mode) { return mode == SEG_PLAIN ? 1 : mode == SEG_UP2 ? 4 : 9; }

struct GemmSeg {
  int mode;     // SEG_*
  int tmap;     // index of the first A tensor map of this segment
  int cblocks;  // 64-wide K blocks (per tap for conv segments)
};

struct GemmParams {
  int M, N;       // rows, accumulator columns (for GEGLU: 2x the output width)
  int num_kb;     // total number of 64-wide K blocks over all segments
  int nseg;
  GemmSeg seg[3];
  // conv tile geometry: a 128-row M tile is a (tn images) x (th rows) x (tw cols) box of the OUTPUT grid
  int tw, th, tn, tiles_x, tiles_y;
  int num_m_tiles, num_n_tiles;  // persistent tile loop: tile = m_tile * num_n_tiles + n_tile
  // epilogue
  void* out;
  int ldo;
  const float* bias;  // [N] (GEGLU: packed like the weight rows) or nullptr
  const void* res;    // residual [M, ldr] or nullptr (may alias out)
  int ldr;
  int geglu;          // 1: out[m, j] = (acc[j] + b[j]) * gelu(acc[BN/2 + j] + b[BN/2 + j]) per tile
  int act;            // 1: out = gelu_erf(acc + b) (diffusers FeedForward(activation_fn="gelu"), the stage-1 prior);
                      // 2: out = silu(acc + b) (TimestepEmbedding.act); applied before the residual; not with geglu
  int epi_tma;        // 1: vectorised epilogue through the staging buffers + TMA (maps.o / maps.r valid)
  // stream-K: the (tile, k-block) iteration space is cut into gridDim.x equal contiguous ranges, so a GEMM whose
  // tile count does not fill the SMs evenly still keeps every tensor core busy.  A CTA whose range starts inside a
  // tile dumps that partial accumulator (fp32) to sk_ws[blockIdx.x] and raises sk_flags[blockIdx.x]; the CTA that
  // owns the tile's first k-block adds the partials of the following CTAs (fixed order => deterministic) and runs
  // the normal epilogue.  Requires all CTAs co-resident (grid <= #SMs, 1 CTA/SM).
  int sk;
  float* sk_ws;         // [gridDim.x][128 * BN]
  unsigned* sk_flags;   // [gridDim.x], zero at rest
  // LayerNorm folded around the GEMM (attention.py:412,429,435 / motion_module.py:226,232 precede every q/k/v and
  // feed-forward projection):  LN(x) W^T = rstd * (x (W*gamma)^T - mean * u) + c,  u[n] = sum_k (W*gamma)[n,k],
  // c[f][n] = sum_k (beta[k] + pe[f][k]) W[n,k] + bias[n].  The folded weight rows are additionally CENTRED
  // (Wf[n,:] = W[n,:]*gamma - mean_k(W[n,:]*gamma)), so x Wf^T already equals x (W*gamma)^T - mean*u and the epilogue
  // is one FMA per element: out = rstd * acc + c.  The GEMM that PRODUCES x emits per-row partial (sum, sum of
  // squares) of its rounded outputs, one float2 per (column-tile half, row): stats_out[part][M]; the GEMM that
  // CONSUMES x sums the parts (only rstd is needed).
  float2* stats_out;       // producer: [2 * num_n_tiles][M], or nullptr
  const float2* stats_in;  // consumer: [stats_parts][M], or nullptr
  int stats_parts;
  int ln_K;                // row length the statistics cover (= K of the consumer)
  const float* ln_c;       // [ln_frames][accumulator columns]
  int ln_frames, ln_rows_per_frame;
  float ln_eps;
  // GroupNorm statistics of the OUTPUT, emitted by the epilogue (GN template flag; resnet.py:185,194 / attention.py:328 /
  // motion_module.py:162 / unet.py:455 consume them): per (image = gn_hw consecutive rows, chunk of GN_CHUNK = 10
  // consecutive channels) the sum and the sum of squares of the rounded output values, accumulated into 128-bit fixed
  // point (two u64 halves each; integer atomics => the result does not depend on the arrival order, so graph replays
  // stay bit-identical): gn_acc[(img * N / 10 + chunk) * 4 + {0: sum hi, 1: sum lo, 2: sumsq hi, 3: sumsq lo}].
  // The consumer (gn_apply_kernel) folds the chunks of a group (any width that is a multiple of 10, over one or all
  // frames, over the two tensors of an un-materialised skip concat) into mean / rstd.  Zeroed once per step.
  unsigned long long* gn_acc;
  int gn_hw;
  // SEG_UP2: output parity class of this launch; out4d: the output tile is stored through a 4-D map (cols, x, y, image)
  int up_py, up_px, out4d;
};
constexpr int GN_CHUNK = 10;
// value v (fp32) -> fixed point v * 2^40 split into (hi = floor(v * 2^-8), lo = remainder < 2^48); exact for |v| >= 2^-16
// (24-bit significand), |error| < 2^-40 below; hi fits 64 bits for every finite fp32 sum of <= 2^20 fp16 values
__device__ __forceinline__ void gn_fixed_split(float v, unsigned long long& hi, unsigned long long& lo) {
  const double d = (double)v * 1099511627776.0;                         // 2^40
  const long long h = __double2ll_rd(d * 3.552713678800501e-15);        // 2^-48
  hi = (unsigned long long)h;
  lo = (unsigned long long)__double2ll_rn(d - (double)h * 281474976710656.0);  // 2^48
}
__host__ __device__ __forceinline__ double gn_fixed_join(unsigned long long hi, unsigned long long lo) {
  return ((double)(long long)hi * 281474976710656.0 + (double)lo) * 9.094947017729282e-13;  // 2^-40
}

// Work decomposition shared by the three warp roles (they must enumerate identical sequences).
struct GemmWork {
  int sk, num_tiles, num_kb, step, tile;
  int nn, step_m, step_n, cur_m, cur_n;  // (m, n) of `tile`, advanced without a division per tile (round-robin mode)
  long long u, u_end;
  __device__ GemmWork(const GemmParams& p, int bid, int nblk) {
    sk = p.sk;
    num_tiles = p.num_m_tiles * p.num_n_tiles;
    num_kb = p.num_kb;
    step = nblk;
    tile = bid;
    nn = p.num_n_tiles;
    step_m = nblk / nn;
    step_n = nblk - step_m * nn;
    cur_m = bid / nn;
    cur_n = bid - cur_m * nn;
    const long long U = (long long)num_tiles * num_kb;
    u = U * bid / nblk;
    u_end = U * (bid + 1) / nblk;
  }
  // next work item: tile index t = mt * num_n_tiles + nt, k-block range [kb0, kb1)
  __device__ bool next(int& t, int& kb0, int& kb1, int& mt, int& nt) {
    if (!sk) {
      if (tile >= num_tiles) return false;
      t = tile;
      mt = cur_m;
      nt = cur_n;
      kb0 = 0;
      kb1 = num_kb;
      tile += step;
      cur_m += step_m;
      cur_n += step_n;
      if (cur_n >= nn) {
        cur_n -= nn;
        ++cur_m;
      }
      return true;
    }
    if (u >= u_end) return false;
    t = (int)(u / num_kb);
    mt = t / nn;
    nt = t - mt * nn;
    kb0 = (int)(u - (long long)t * num_kb);
    const long long len = (u_end - u) < (long long)(num_kb - kb0) ? (u_end - u) : (long long)(num_kb - kb0);
    kb1 = kb0 + (int)len;
    u += len;
    return true;
  }
  __device__ bool next(int& t, int& kb0, int& kb1) {
    int mt, nt;
    return next(t, kb0, kb1, mt, nt);
  }
};

struct GemmMaps {
  CUtensorMap a[4];
  CUtensorMap b;
  CUtensorMap o;  // output   [M, n_out]: box (columns of one epilogue warp group, 128 rows), TMA store
  CUtensorMap r;  // residual [M, n_out]: same box, TMA load into the staging buffer
};

// PAIR = true: CTA-pair kernel (cluster of 2, tcgen05 cta_group::2).  The pair computes a 256 x BN tile: CTA r holds A
// rows [128r, 128r+128) and B rows [r*BN/2, (r+1)*BN/2) in its own shared memory and its 128 accumulator rows in its own
// TMEM.  Per MMA the SM then ingests 16 KB + BN*64 B instead of 16 KB + BN*128 B: the L2 -> SM port (~64 B/clk), not
// the tensor pipe, bounds the single-CTA kernel (128x160: 115 B/clk at full MMA rate; pair 256x160: 83 B/clk).
#ifndef RCDM_GEMM_STAGES64
#define RCDM_GEMM_STAGES64 6
#endif
template <int BN, bool PAIR = false> struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  // (RCDM_GEMM_EXPERIMENT == 11: one stage less, to measure how far the main loop is bound by stages / load latency)
  // (192-wide tiles - single CTA only - are used where they save a whole wave of tiles, see gemm_plain_bn: 3 stages)
  static constexpr int STAGES = (PAIR ? (BN <= 64 ? 8 : BN <= 128 ? 6 : 5)
                                      : ((BN <= 64) ? RCDM_GEMM_STAGES64 : (BN <= 128 ? 5 : (BN <= 160 ? 4 : 3)))) -
                                (RCDM_GEMM_EXPERIMENT == 11 ? 1 : 0);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NG = RCDM_EPI_GROUPS;   // column groups; 4 * NG epilogue warps
  static constexpr int QW = BN / NG;           // accumulator columns per epilogue warp
  static constexpr int NPROD = RCDM_GEMM_PRODUCERS;
  // producer 0, MMA, 4 * NG epilogue warps, store warp, producers 1..NPROD-1
  static constexpr int THREADS = 64 + 128 * NG + 32 + 32 * (NPROD - 1);
  static constexpr int STORE_WARP = 2 + 4 * NG;
  // staging buffer = the 128 x BN 16-bit output tile as NG dense [128 rows][QW] parts (TMA box layout); two buffers
  // (tile parity) so the residual of tile i+1 loads while tile i is written and stored
  static constexpr int PART_BYTES = 128 * QW * 2;
  static constexpr int STG_BYTES = NG * PART_BYTES;
  static constexpr int STAGING_BYTES = 2 * STG_BYTES;
  static constexpr int ACC_STRIDE = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // TMEM columns between the 2 accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  // BN fp32 epilogue values (bias, or the folded-LayerNorm vector c) per TMEM lane quarter: every epilogue warp owns a
  // private slice (its quarter's row, its column group), so no shared-memory word is written by one warp and read by
  // another (compute-sanitizer racecheck clean) and no tile-parity double buffer is needed
  static constexpr int BIAS_BYTES = 4 * BN * 4;
  // the dynamic shared memory window is declared __align__(1024) (128B-swizzle atoms), so there is no alignment slack
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256 /*barriers*/ + BIAS_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};

// Persistent, warp-specialised: grid = min(#tiles, #SMs), one CTA per SM.  Tiles are visited round-robin
// (n fastest, so CTAs running concurrently share A rows in L2).  The TMA producer runs ahead across tile
// boundaries; two TMEM accumulators let the MMA warp start tile i+1 while the epilogue warps drain tile i.
// 12 warps = 3 per SM sub-partition (16384 registers each) => at most 168 registers per thread.
// ACT: compile-time switch for the epilogue activation (p.act).  The activation-free instantiation is the one every
// UNet GEMM runs: a run-time branch inside `finish8` cost the epilogue-bound small-K GEMMs 15-19 % (measured), so the
// stage-1 prior's GELU / SiLU epilogues get their own instantiation.
template <typename T, int BN, bool PAIR, bool ACT = false, bool GN = false>
__global__ void __launch_bounds__(GemmCfg<BN, PAIR>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const GemmParams p) {
  using Cfg = GemmCfg<BN, PAIR>;
  // pair mode: the work decomposition runs over PAIRS (p.num_m_tiles counts 256-row tile pairs); this CTA's rows are
  // m_tile = 2 * pair_tile + rank
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int wid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // worker (CTA or pair) index
  const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int sk_stride = PAIR ? 2 : 1;                                     // stream-K slots: one per CTA
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {  // never observed; a misaligned window would corrupt the swizzled tiles
    if (threadIdx.x == 0) printf("rcdm: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
  uint64_t* res_full = bars + 2 * STAGES + 4;        // [2] residual tile landed in staging[b]
  uint64_t* stg_full = bars + 2 * STAGES + 6;        // [2] every epilogue warp has written its slice of staging[b]
  uint64_t* stg_free = bars + 2 * STAGES + 8;        // [2] the TMA store has finished reading staging[b]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 10);
  static_assert((2 * STAGES + 10) * 8 + 4 <= 256, "barrier block overflows its 256 bytes");
  float* bias_sm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [4 lane quarters][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.a[0]);
    if (p.epi_tma) {
      tma_prefetch_desc(&maps.o);
      if (p.res) tma_prefetch_desc(&maps.r);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full_bar[i], 1);
        mbar_init(&tmem_empty_bar[i], (PAIR ? 8 : 4) * Cfg::NG);  // one arrival per epilogue warp (pair: of both CTAs)
        mbar_init(&res_full[i], 1);
        mbar_init(&stg_full[i], 4 * Cfg::NG);  // one arrival per epilogue warp of THIS CTA
        mbar_init(&stg_free[i], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (PAIR) tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  __syncthreads();  // (also in a pair: the CTA-scope barrier is what orders the TMEM-slot write for racecheck)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0 || warp > Cfg::STORE_WARP) {
    // =================================== TMA producers ===================================
    // every producer walks the same (tile, k-block) sequence; k-block number g of this CTA belongs to producer g % NPROD
    if (elect_one()) {
      const int pid = warp == 0 ? 0 : warp - Cfg::STORE_WARP;
      int turn = 0;  // g % NPROD
      int stage = 0;
      uint32_t phase = 0;
      GemmWork work(p, wid, nworkers);
      int tile, kb0, kb1, mt_, n_tile;
      int pit = -1;  // work items so far (trace)
      while (work.next(tile, kb0, kb1, mt_, n_tile)) {
        ++pit;
        const int m_tile = PAIR ? 2 * mt_ + (int)rank : mt_;
        int t = m_tile;  // conv tile origin (only used by conv segments)
        const int tx = t % p.tiles_x;
        t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int tb = t / p.tiles_y;
        const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tb * p.tn;
        // decode kb0 -> (segment, tap, channel block)
        int s = 0, tap = 0, c = 0;
        {
          int rem = kb0;
          for (; s < p.nseg - 1; ++s) {
            const int n = seg_taps(p.seg[s].mode) * p.seg[s].cblocks;
            if (rem < n) break;
            rem -= n;
          }
          tap = rem / p.seg[s].cblocks;
          c = rem - tap * p.seg[s].cblocks;
        }
        GemmSeg sg = p.seg[s];
#if RCDM_GEMM_EXPERIMENT == 10
        // experiment: L2 prefetch of the NEXT tile's A rows (plain single-segment GEMMs, no stream-K)
        if (!p.sk && p.nseg == 1 && sg.mode == SEG_PLAIN && tile + work.step < work.num_tiles) {
          const int nt = tile + work.step;
          const int nm = PAIR ? 2 * (nt / p.num_n_tiles) + (int)rank : nt / p.num_n_tiles;
          if (nm != m_tile)
            for (int c2 = 0; c2 < sg.cblocks; ++c2) tma_prefetch_2d(&maps.a[sg.tmap], c2 * 64, nm * 128);
        }
#endif
        for (int kb = kb0; kb < kb1; ++kb) {
          int mi = sg.tmap, dx = 0, dy = 0;
          if (sg.mode == SEG_CONV3) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          } else if (sg.mode == SEG_CONV3S2) {
            const int ky = tap / 3, kx = tap % 3;
            const int py = (ky + 1) & 1, px = (kx + 1) & 1;  // parity of (2y + ky - 1)
            dy = (ky == 0) ? -1 : 0;
            dx = (kx == 0) ? -1 : 0;
            mi = sg.tmap + py * 2 + px;
          } else if (sg.mode == SEG_UP2) {
            dy = (tap >> 1) - 1 + p.up_py;
            dx = (tap & 1) - 1 + p.up_px;
          } else if (sg.mode == SEG_CONV3S2A) {
            const int ky = tap / 3, kx = tap % 3;
            dy = ky >> 1;
            dx = kx >> 1;
            mi = sg.tmap + (ky & 1) * 2 + (kx & 1);
          }
          if (turn == pid) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (kb == kb0 || kb == kb0 + 1) GEMM_STAMP(pit, pid);              // first k-block of this producer: slot free
          void* sa = smem_a + stage * Cfg::A_BYTES;
          void* sb = smem_b + stage * Cfg::B_BYTES;
          if constexpr (PAIR) {
            // both CTAs' loads complete on the LEADER's full barrier (the leader alone issues the MMA)
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            if (sg.mode == SEG_PLAIN)
              tma_load_2d_2sm(sa, &maps.a[mi], &full_bar[stage], c * 64, m_tile * 128);
            else
              tma_load_4d_2sm(sa, &maps.a[mi], &full_bar[stage], c * 64, x0 + dx, y0 + dy, n0);
            tma_load_2d_2sm(sb, &maps.b, &full_bar[stage], kb * 64, n_tile * BN + (int)rank * (BN / 2));
          } else {
#if RCDM_GEMM_EXPERIMENT == 3 || RCDM_GEMM_EXPERIMENT == 4 || RCDM_GEMM_EXPERIMENT == 7 || RCDM_GEMM_EXPERIMENT == 9
#if RCDM_GEMM_EXPERIMENT == 7
            const bool load_a = true, load_b = tile == wid;
#else
            const bool load_a = kb == kb0, load_b = (RCDM_GEMM_EXPERIMENT == 3) || kb == kb0;
#endif
            mbar_expect_tx(&full_bar[stage], (load_a ? Cfg::A_BYTES : 0) + (load_b ? Cfg::B_BYTES : 0));
            if (load_a) {
              if (sg.mode == SEG_PLAIN) tma_load_2d(sa, &maps.a[mi], &full_bar[stage], c * 64, m_tile * 128);
              else tma_load_4d(sa, &maps.a[mi], &full_bar[stage], c * 64, x0 + dx, y0 + dy, n0);
            }
            if (load_b) tma_load_2d(sb, &maps.b, &full_bar[stage], kb * 64, n_tile * BN);
#else
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if (sg.mode == SEG_PLAIN)
              tma_load_2d(sa, &maps.a[mi], &full_bar[stage], c * 64, m_tile * 128);
            else
              tma_load_4d(sa, &maps.a[mi], &full_bar[stage], c * 64, x0 + dx, y0 + dy, n0);
            tma_load_2d(sb, &maps.b, &full_bar[stage], kb * 64, n_tile * BN);
#endif
          }
          if (kb >= kb1 - 2) GEMM_STAMP(pit, 2 + pid);                       // last k-block of this producer: issued
          }
          if (++turn == Cfg::NPROD) turn = 0;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          if (++c == sg.cblocks) {  // advance (segment, tap, channel block)
            c = 0;
            if (++tap == seg_taps(sg.mode)) {
              tap = 0;
              if (s + 1 < p.nseg) sg = p.seg[++s];
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =================================== MMA issuer ===================================
    if ((!PAIR || rank == 0) && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(DT<T>::umma_fmt, PAIR ? 256 : 128, BN, 0, 0);
      // K-major, 128B swizzle: rows at 128 B, 8-row groups at 1024 B; +32 B per 16-element K step (address field is
      // bytes >> 4).  (Unrolling this loop over the stages with compile-time descriptors was measured: no gain for
      // large K - the issue thread is not the limiter - and register spills in the epilogue for small K.)
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      GemmWork work(p, wid, nworkers);
      int tile, kb0, kb1;
      for (; work.next(tile, kb0, kb1); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        GEMM_STAMP(it, 4);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (kb == kb0) GEMM_STAMP(it, 5);
          if (kb == kb1 - 1) GEMM_STAMP(it, 6);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
            const uint64_t bd = umma_smem_desc(b_addr + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
            if constexpr (PAIR) umma_f16_ss_2cta(d_tmem, ad, bd, idesc, ((kb - kb0) | k) != 0);
            else umma_f16_ss(d_tmem, ad, bd, idesc, ((kb - kb0) | k) != 0);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs have read it
          if constexpr (PAIR) umma_commit_2cta(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (PAIR) umma_commit_2cta(&tmem_full_bar[acc]);
        else umma_commit(&tmem_full_bar[acc]);
        GEMM_STAMP(it, 7);
      }
    }
  } else if (warp < 2 + 4 * Cfg::NG) {
    // =================================== epilogue (4 * NG warps) ===================================
    // warp -> (q, cg): q = TMEM lane quarter (rows q*32..+31), cg = which column group of the tile.
    // thread <-> row: TMEM -> registers, scale * acc + vector (smem broadcast), GEGLU, + residual (packed half2 add ==
    // fp32 add + one rounding) -> this row's slice of the staging buffer; then one thread issues the TMA stores.
    constexpr int NG = Cfg::NG;
    constexpr int QW = Cfg::QW;          // accumulator columns per warp
    constexpr int TH = BN / 2;           // GEGLU: gate columns start at TH
    constexpr int CW = (QW % 16 == 0) ? 16 : 8;  // TMEM columns per tcgen05.ld
    constexpr int NCH = QW / CW;
    constexpr int EPI_THREADS = 128 * NG;
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    T* out = reinterpret_cast<T*>(p.out);
#if RCDM_GEMM_EXPERIMENT == 6
    const T* res = nullptr;
#else
    const T* res = reinterpret_cast<const T*>(p.res);
#endif
    const bool vec_ok = p.epi_tma != 0;
    const int n_total = p.geglu ? p.N / 2 : p.N;
    const int wcols = p.geglu ? QW / 2 : QW;       // output columns this warp produces per tile
    const int tile_cols = p.geglu ? TH : BN;       // output columns per tile
    auto ld_chunk = [&](uint32_t addr, uint32_t* r) {
      if constexpr (CW == 16) tmem_ld16(addr, r);
      else tmem_ld8(addr, r);
    };
    auto st_chunk = [&](uint32_t addr, const uint32_t* r) {
      if constexpr (CW == 16) tmem_st16(addr, r);
      else tmem_st8(addr, r);
    };
    auto epi_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); };
    int it = 0;
    int ot = 0;  // output-producing work items so far (stream-K partial dumps do not count)
    GemmWork work(p, wid, nworkers);
    int tile, kb0, kb1, mt_, n_tile;
    // "accumulator drained": in a pair every epilogue thread of both CTAs arrives on the LEADER's barrier
    // one arrival per warp (after a warp sync), not per thread: 256 serialised arrivals on one mbarrier per tile were
    // a measurable part of the per-tile latency chain  epilogue -> tmem_empty -> next MMA
    auto release_acc = [&](int acc) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
    };
    for (; work.next(tile, kb0, kb1, mt_, n_tile); ++it) {
      const int m_tile = PAIR ? 2 * mt_ + (int)rank : mt_;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + acc * Cfg::ACC_STRIDE + (uint32_t(q * 32) << 16);
      const int m_warp = m_tile * 128 + q * 32;
#if RCDM_GEMM_EXPERIMENT == 5 || RCDM_GEMM_EXPERIMENT == 9
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      release_acc(acc);
      continue;
#endif
      if (kb0 > 0) {
        // ---- stream-K partial: this CTA's range began inside the tile -> dump the fp32 accumulator, raise the flag
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        // workspace layout = the warps' own access order: [slot][q][cg][chunk][CW/4 x uint4][lane] -> every
        // store / load instruction of a warp covers 512 contiguous bytes
        uint4* ws = reinterpret_cast<uint4*>(p.sk_ws) + (size_t)blockIdx.x * (128 * BN / 4) +
                    (size_t)(q * NG + cg) * (NCH * (CW / 4) * 32) + lane;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          uint32_t r[CW];
          ld_chunk(taddr + cg * QW + ch * CW, r);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < CW / 4; ++g)
            __stcg(ws + (ch * (CW / 4) + g) * 32, make_uint4(r[g * 4], r[g * 4 + 1], r[g * 4 + 2], r[g * 4 + 3]));
        }
        release_acc(acc);
        epi_bar();  // all epilogue warps have stored their slice
        if (warp == 2 && lane == 0) {
          __threadfence();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sk_flags + blockIdx.x), "r"(1u) : "memory");
        }
        continue;
      }
      // stream-K head: the CTAs after this one hold the rest of the tile's K range
      int ncontrib = 0;
      if (p.sk && kb1 < p.num_kb) {
        const long long U = (long long)work.num_tiles * p.num_kb, tile_end = (long long)(tile + 1) * p.num_kb;
        int j = wid + 1;
        while (j < nworkers && U * j / nworkers < tile_end) ++j;
        ncontrib = j - wid - 1;
      }
      // Fix-up pre-pass (at most one tile per CTA): wait for the contributors' flags, add their partials into the
      // TMEM accumulator (all NCH*4 16-byte loads of a contributor in flight at once), re-arm the flags; the normal
      // epilogue below then runs unchanged.
      auto sk_fixup = [&]() {
        if (ncontrib == 0) return;
        if (lane == 0) {
          for (int j = 1; j <= ncontrib; ++j) {
            unsigned spins = 0, seen = 0;
            while (true) {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.sk_flags + blockIdx.x + j * sk_stride) : "memory");
              if (seen) break;
              if (++spins > (1u << 26)) {
                printf("rcdm: stream-K partial never arrived (block %d waits for %d)\n", blockIdx.x, blockIdx.x + j * sk_stride);
                __trap();
              }
            }
          }
        }
        __syncwarp();
        for (int j = 1; j <= ncontrib; ++j) {
          const uint4* src = reinterpret_cast<const uint4*>(p.sk_ws) + (size_t)(blockIdx.x + j * sk_stride) * (128 * BN / 4) +
                             (size_t)(q * NG + cg) * (NCH * (CW / 4) * 32) + lane;
          uint4 pr[NCH * (CW / 4)];  // every 16-byte load of this contributor in flight at once
#pragma unroll
          for (int i = 0; i < NCH * (CW / 4); ++i) pr[i] = __ldcg(src + i * 32);
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            uint32_t r[CW];
            ld_chunk(taddr + cg * QW + ch * CW, r);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < CW / 4; ++g) {
              const uint4 v = pr[ch * (CW / 4) + g];
              r[g * 4] = __float_as_uint(__uint_as_float(r[g * 4]) + __uint_as_float(v.x));
              r[g * 4 + 1] = __float_as_uint(__uint_as_float(r[g * 4 + 1]) + __uint_as_float(v.y));
              r[g * 4 + 2] = __float_as_uint(__uint_as_float(r[g * 4 + 2]) + __uint_as_float(v.z));
              r[g * 4 + 3] = __float_as_uint(__uint_as_float(r[g * 4 + 3]) + __uint_as_float(v.w));
            }
            st_chunk(taddr + cg * QW + ch * CW, r);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        epi_bar();  // all partials consumed; other warps' columns are final
        tc_fence_after();
        if (warp == 2 && lane == 0)
          for (int j = 1; j <= ncontrib; ++j) p.sk_flags[blockIdx.x + j * sk_stride] = 0;  // re-arm for the next launch
      };
      if (ncontrib > 0) {  // stream-K head (rare): fold the partials in first, while no prefetched residual is live
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        sk_fixup();
      }
      // folded LayerNorm (consumer side): statistics of this thread's row, summed over the producer's column parts;
      // out = ln_a * acc + c[frame][n]   (ln_a = rstd of the row; the weights are centred, see GemmParams)
      const bool ln = p.stats_in != nullptr;
      float ln_a = 1.f;
      int ln_f = 0, ln_f0 = 0;
      // c from shared memory when the whole tile lies in one frame (always, except the 8x8 level of the temporal
      // projections where a 128-row tile spans two frames: those read c through L1)
      const bool ln_smem = p.ln_frames == 1 || (p.ln_rows_per_frame % 128) == 0;
      constexpr int ST_PRE = 4;  // statistic parts fetched ahead (C = 320: 4 parts, 640: 8, 1280: 16); more would spill
      float2 st_pre[ST_PRE];
      float st_xs = 0.f, st_xss = 0.f;  // parts beyond the prefetched ones (wide producers: N = 2048 of the prior emits 32)
      const int ln_m = min(m_warp + lane, p.M - 1);
      if (ln) {
#pragma unroll
        for (int pp = 0; pp < ST_PRE; ++pp)
          if (pp < p.stats_parts) st_pre[pp] = __ldg(&p.stats_in[(size_t)pp * p.M + ln_m]);
        // summed here, ahead of the accumulator wait as well (two live registers instead of a load chain of
        // (parts - 4) / unroll L2 round trips between the wait and the first output column)
#pragma unroll 4
        for (int pp = ST_PRE; pp < p.stats_parts; ++pp) {
          const float2 v = __ldg(&p.stats_in[(size_t)pp * p.M + ln_m]);
          st_xs += v.x;
          st_xss += v.y;
        }
        if (p.ln_frames > 1) {
          ln_f = (ln_m / p.ln_rows_per_frame) % p.ln_frames;
          ln_f0 = (min(m_tile * 128, p.M - 1) / p.ln_rows_per_frame) % p.ln_frames;
        }
      }
      // consumed only after the accumulator wait, so the L2 latency of the loads above hides behind the main loop
      auto ln_finish = [&]() {
        float sx = 0.f, sxx = 0.f;
#pragma unroll
        for (int pp = 0; pp < ST_PRE; ++pp)
          if (pp < p.stats_parts) {
            sx += st_pre[pp].x;
            sxx += st_pre[pp].y;
          }
        sx += st_xs;
        sxx += st_xss;
        const float inv_k = 1.0f / (float)p.ln_K;
        const float mean = sx * inv_k;
        const float var = fmaxf(sxx * inv_k - mean * mean, 0.f);
        ln_a = rsqrtf(var + p.ln_eps);
      };
      const float* ln_crow = p.ln_c + (size_t)ln_f * p.N;  // global fallback row of c
      if (vec_ok) {
        const int n_warp = n_tile * tile_cols + cg * wcols;  // first output column of this warp
        const int sb = ot & 1;                               // staging buffer of this tile
        // ---- epilogue vector of this warp's columns (bias, or LN c): fetched into registers BEFORE the accumulator
        // wait (latency hidden), parked in this warp's private smem slice AFTER it (broadcast reads in the column loop)
        constexpr int PV = (QW + 31) / 32;
        float pre0[PV];
        {
          const float* v0 = ln ? p.ln_c + (size_t)ln_f0 * p.N : p.bias;
#pragma unroll
          for (int j = 0; j < PV; ++j) {
            const int i = lane + j * 32;
            pre0[j] = 0.f;
            if (i < QW && v0) {
              if (!p.geglu) {
                if (n_warp + i < n_total) pre0[j] = __ldg(v0 + n_warp + i);
              } else {  // packed GEGLU vectors: tile-local [h (TH) | gate (TH)]; this warp: h/gate cols cg*wcols..
                pre0[j] = __ldg(v0 + n_tile * BN + (i < wcols ? cg * wcols + i : TH + cg * wcols + (i - wcols)));
              }
            }
          }
        }
        if (warp == 2 && lane == 0) GEMM_STAMP(it, 8);
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        if (warp == 2 && lane == 0) GEMM_STAMP(it, 9);
        tc_fence_after();
        if (ln) ln_finish();
        float* bsm = bias_sm + q * BN + cg * QW;  // bias, or c of the tile's frame in LN mode (this warp's slice)
#pragma unroll
        for (int j = 0; j < PV; ++j) {
          const int i = lane + j * 32;
          if (i < QW) bsm[i] = pre0[j];
        }
        __syncwarp();
        // staging[sb] is ours to write: the residual tile has landed in it (the store warp loads it only after the store
        // of the tile that used the buffer before has been read), or - no residual - that store has been read
        if (res) mbar_wait(&res_full[sb], (ot >> 1) & 1);
        else mbar_wait(&stg_free[sb], ((ot >> 1) & 1) ^ 1);
        if (warp == 2 && lane == 0) GEMM_STAMP(it, 10);
        // ---- thread <-> row: v = scale * acc + vec (+GEGLU), rounded to 16 bits, (+ residual, rounded again like the
        // reference's `x + attn(...)` on 16-bit tensors), written to this row's slice of the staging buffer
        const float scale = ln_a;
        uint8_t* srow = staging + sb * Cfg::STG_BYTES + cg * Cfg::PART_BYTES + (size_t)(q * 32 + lane) * (wcols * 2);
        float st_s = 0.f, st_ss = 0.f;  // row statistics of the final values (folded-LayerNorm producer)
        constexpr int GNC = GN ? QW / GN_CHUNK : 1;  // GroupNorm chunks of this warp's columns
        float gn_s[GNC], gn_ss[GNC];
#pragma unroll
        for (int i = 0; i < GNC; ++i) gn_s[i] = gn_ss[i] = 0.f;
        using T2 = typename DT<T>::T2;
        // 8 outputs at columns c.. of this warp's slice; returns them packed.  rv_pre: the residual values of these 8
        // columns when the caller fetched them ahead (see the software pipeline below), else read from the staging row here.
        // RESC / STC (std::bool_constant): residual add / row statistics, compile-time so that the unrolled column loop is
        // straight-line code (ncu, K = 320 GEMMs: with run-time tests per 8-column step a quarter of the epilogue warps'
        // samples were instruction-fetch stalls behind the ~60 branches per tile, another tenth branch resolution).
        // Columns beyond N need no test: their accumulators, epilogue-vector entries and (TMA zero-filled) residual values
        // are all zero, so they add nothing to the statistics; a warp whose columns lie wholly beyond N passes STC = false.
        auto finish8 = [&](auto RESC, auto STC, float* v, int c, const uint4* rv_pre = nullptr) -> uint4 {
          if constexpr (ACT) {
            if (p.act == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = gelu_erf_f(v[i]);
            } else if (p.act == 2) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
            }
          }
          uint4 pk = pack8<T>(v);
          uint4* dst = reinterpret_cast<uint4*>(srow + c * 2);
          if constexpr (decltype(RESC)::value) {
            const uint4 rv = rv_pre ? *rv_pre : *dst;
            T2* a2 = reinterpret_cast<T2*>(&pk);
            const T2* b2 = reinterpret_cast<const T2*>(&rv);
#pragma unroll
            for (int i = 0; i < 4; ++i) a2[i] = __hadd2(a2[i], b2[i]);
          }
          if constexpr (decltype(STC)::value) {
            float f8[8];
            unpack8<T>(pk, f8);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              st_s += f8[i];
              st_ss = fmaf(f8[i], f8[i], st_ss);
            }
          }
          *dst = pk;
          return pk;
        };
        const bool do_stats = p.stats_out != nullptr && n_warp < n_total;
        // run-time flags -> one of the four instantiations of `fn` (once per tile, outside the column loop)
        auto dispatch = [&](auto&& fn) {
          if (res) {
            if (do_stats) fn(std::true_type{}, std::true_type{});
            else fn(std::true_type{}, std::false_type{});
          } else {
            if (do_stats) fn(std::false_type{}, std::true_type{});
            else fn(std::false_type{}, std::false_type{});
          }
        };
        // GroupNorm chunk sums of 8 final values at columns c.. ; called only where c is a compile-time constant after
        // unrolling (the accumulators must stay in registers)
        auto gn_add8 = [&](const uint4& pk, int c) {
          if constexpr (GN) {
            float f8[8];
            unpack8<T>(pk, f8);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              gn_s[(c + i) / GN_CHUNK] += f8[i];
              gn_ss[(c + i) / GN_CHUNK] = fmaf(f8[i], f8[i], gn_ss[(c + i) / GN_CHUNK]);
            }
          }
        };
        if (ln && !ln_smem) {
          // rare: a 128-row tile of a temporal projection spans several frames (8x8 level, tiny test configs): the
          // per-frame vector c comes through L1 instead of shared memory
          if (!p.geglu) {
#pragma unroll 1
            for (int c = 0; c < QW; c += 8) {
              uint32_t r[8];
              tmem_ld8(taddr + cg * QW + c, r);
              tmem_wait_ld();
              const int col = n_tile * BN + cg * QW + c;
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                v[i] = fmaf(__uint_as_float(r[i]), scale, (col + i < p.N) ? __ldg(ln_crow + col + i) : 0.f);
              dispatch([&](auto RESC, auto STC) { finish8(RESC, STC, v, c); });
            }
          } else {
#pragma unroll 1
            for (int c = 0; c < QW / 2; c += 8) {
              uint32_t rh[8], rg[8];
              tmem_ld8(taddr + cg * (QW / 2) + c, rh);
              tmem_ld8(taddr + TH + cg * (QW / 2) + c, rg);
              tmem_wait_ld();
              const int colh = n_tile * BN + cg * (QW / 2) + c;
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                v[i] = fmaf(__uint_as_float(rh[i]), scale, __ldg(ln_crow + colh + i)) *
                       gelu_erf_f(fmaf(__uint_as_float(rg[i]), scale, __ldg(ln_crow + colh + TH + i)));
              finish8(std::false_type{}, std::false_type{}, v, c);  // (GEGLU: no residual, no statistics)
            }
          }
        } else if (!p.geglu) {
          // Software pipeline over steps of 8 columns: the TMEM load of chunk c+1 is in flight while chunk c is processed,
          // and the shared-memory operands of step s+1 (8 epilogue-vector values, 8 residual values) are requested BEFORE
          // step s is computed and stored.  (ncu, K = 320 GEMMs: with the loads issued where they are used, every FFMA
          // waited ~30 cycles on its LDS - the compiler may not hoist a shared load above the preceding staging store -
          // and the 2 epilogue warps per scheduler could not hide it: 0.2 instructions per cycle.)
          dispatch([&](auto RESC, auto STC) {
            constexpr int SPC = CW / 8;     // steps per TMEM chunk
            constexpr int NS = QW / 8;      // steps per tile
            uint32_t r[2][CW];
            float4 bq[2][2];
            uint4 rq[2];
            auto fetch = [&](int st, int slot) {
              bq[slot][0] = *reinterpret_cast<const float4*>(bsm + st * 8);  // smem broadcast
              bq[slot][1] = *reinterpret_cast<const float4*>(bsm + st * 8 + 4);
              if constexpr (decltype(RESC)::value) rq[slot] = *reinterpret_cast<const uint4*>(srow + st * 16);
            };
            ld_chunk(taddr + cg * QW, r[0]);
            fetch(0, 0);
#pragma unroll
            for (int st = 0; st < NS; ++st) {
              const int ci = st / SPC, g = st % SPC;
              if (g == 0) {
                tmem_wait_ld();
                if (ci + 1 < NCH) ld_chunk(taddr + cg * QW + (ci + 1) * CW, r[(ci + 1) & 1]);
              }
              if (st + 1 < NS) fetch(st + 1, (st + 1) & 1);
              const uint32_t* rc = r[ci & 1];
              const float4 b0 = bq[st & 1][0], b1 = bq[st & 1][1];
              float v[8];
              v[0] = fmaf(__uint_as_float(rc[g * 8]), scale, b0.x);
              v[1] = fmaf(__uint_as_float(rc[g * 8 + 1]), scale, b0.y);
              v[2] = fmaf(__uint_as_float(rc[g * 8 + 2]), scale, b0.z);
              v[3] = fmaf(__uint_as_float(rc[g * 8 + 3]), scale, b0.w);
              v[4] = fmaf(__uint_as_float(rc[g * 8 + 4]), scale, b1.x);
              v[5] = fmaf(__uint_as_float(rc[g * 8 + 5]), scale, b1.y);
              v[6] = fmaf(__uint_as_float(rc[g * 8 + 6]), scale, b1.z);
              v[7] = fmaf(__uint_as_float(rc[g * 8 + 7]), scale, b1.w);
              gn_add8(finish8(RESC, STC, v, st * 8, &rq[st & 1]), st * 8);
            }
          });
        } else {
          // GEGLU: h columns [cg * QW/2, +QW/2), gate columns TH + the same; outputs QW/2 per warp.  Same software
          // pipeline: the 16 epilogue-vector values of step s+1 are requested before step s is computed.
          constexpr int GW = QW / 2;
          constexpr int GCW = (GW % 16 == 0) ? 16 : 8;
          constexpr int GNCH = GW / GCW;
          constexpr int GSPC = GCW / 8, GNS = GW / 8;
          uint32_t rh[2][GCW], rg[2][GCW];
          float4 bh[2][2], bg[2][2];
          auto ldg_chunk = [&](uint32_t addr, uint32_t* r) {
            if constexpr (GCW == 16) tmem_ld16(addr, r);
            else tmem_ld8(addr, r);
          };
          auto fetch = [&](int st, int slot) {
            bh[slot][0] = *reinterpret_cast<const float4*>(bsm + st * 8);
            bh[slot][1] = *reinterpret_cast<const float4*>(bsm + st * 8 + 4);
            bg[slot][0] = *reinterpret_cast<const float4*>(bsm + wcols + st * 8);
            bg[slot][1] = *reinterpret_cast<const float4*>(bsm + wcols + st * 8 + 4);
          };
          ldg_chunk(taddr + cg * GW, rh[0]);
          ldg_chunk(taddr + TH + cg * GW, rg[0]);
          fetch(0, 0);
#pragma unroll
          for (int st = 0; st < GNS; ++st) {
            const int ci = st / GSPC, g = st % GSPC;
            if (g == 0) {
              tmem_wait_ld();
              if (ci + 1 < GNCH) {
                ldg_chunk(taddr + cg * GW + (ci + 1) * GCW, rh[(ci + 1) & 1]);
                ldg_chunk(taddr + TH + cg * GW + (ci + 1) * GCW, rg[(ci + 1) & 1]);
              }
            }
            if (st + 1 < GNS) fetch(st + 1, (st + 1) & 1);
            const uint32_t* ph = rh[ci & 1];
            const uint32_t* pg = rg[ci & 1];
            const float bhv[8] = {bh[st & 1][0].x, bh[st & 1][0].y, bh[st & 1][0].z, bh[st & 1][0].w,
                                  bh[st & 1][1].x, bh[st & 1][1].y, bh[st & 1][1].z, bh[st & 1][1].w};
            const float bgv[8] = {bg[st & 1][0].x, bg[st & 1][0].y, bg[st & 1][0].z, bg[st & 1][0].w,
                                  bg[st & 1][1].x, bg[st & 1][1].y, bg[st & 1][1].z, bg[st & 1][1].w};
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              v[i] = fmaf(__uint_as_float(ph[g * 8 + i]), scale, bhv[i]) *
                     gelu_erf_f(fmaf(__uint_as_float(pg[g * 8 + i]), scale, bgv[i]));
            finish8(std::false_type{}, std::false_type{}, v, st * 8);
          }
        }
        release_acc(acc);  // accumulator drained: the MMA warp may reuse it
        if (warp == 2 && lane == 0) GEMM_STAMP(it, 11);
        if (p.stats_out && m_warp + lane < p.M)
          p.stats_out[(size_t)(n_tile * NG + cg) * p.M + m_warp + lane] = make_float2(st_s, st_ss);
        if constexpr (GN) {
          // sum the 2 * GNC per-row values over the warp's 32 rows with a reduce-scatter butterfly (fixed tree, 2 * GNC
          // shuffles instead of 5 per value): afterwards lanes 2k and 2k+1 hold value k (k < GNC: sum of chunk k,
          // else sum of squares of chunk k - GNC); lane 2k adds it to the tensor's fixed-point accumulators
          constexpr int NV = 2 * GNC;  // 16 (two column groups) or 8 (four): a power of two <= 32
          static_assert(NV == 16 || NV == 8, "the butterfly below reduces 16 or 8 values over 32 lanes");
          float vals[NV];
#pragma unroll
          for (int i = 0; i < GNC; ++i) {
            vals[i] = gn_s[i];
            vals[GNC + i] = gn_ss[i];
          }
#pragma unroll
          for (int w = NV / 2, m = 16; w >= 1; w >>= 1, m >>= 1) {
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int i = 0; i < w; ++i) {
              const float send = up ? vals[i] : vals[i + w];
              const float keep = up ? vals[i + w] : vals[i];
              vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
          }
          // vals[0] now holds value k = (lane >> SH) & (NV - 1) summed over the lanes that differ in the top log2(NV)
          // bits; the remaining SH low bits are folded with plain butterflies
          constexpr int SH = NV == 16 ? 1 : 2;
          float total = vals[0];
#pragma unroll
          for (int o = 1 << (SH - 1); o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
          if ((lane & ((1 << SH) - 1)) == 0 && m_warp < p.M) {  // (a pair's last 256-row tile may be half outside M)
            const int k = lane >> SH;
            const int img = m_warp / p.gn_hw;
            const int chunk = n_warp / GN_CHUNK + (k % GNC);
            unsigned long long hi, lo;
            gn_fixed_split(total, hi, lo);
            unsigned long long* a = p.gn_acc + ((size_t)img * (p.N / GN_CHUNK) + chunk) * 4 + (k / GNC) * 2;
            atomicAdd(a, hi);
            atomicAdd(a + 1, lo);
          }
        }
        fence_proxy_async_smem();  // this thread's staging writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(&stg_full[sb]);  // the store warp issues the TMA stores when all 4 * NG warps arrived
        ++ot;
      } else {
        // ---- scalar fallback (conv_out: N = 4): thread <-> row, direct stores; only the cg == 0 warps work
        const int m = m_warp + lane;
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        if (cg == 0) {
#pragma unroll 1
          for (int c = 0; c < BN; c += 16) {
            const int n = n_tile * BN + c;
            if (n >= p.N) break;  // warp-uniform
            uint32_t r[16];
            tmem_ld16(taddr + c, r);
            tmem_wait_ld();
            if (m < p.M) {
              for (int i = 0; i < 16; ++i) {
                if (n + i < p.N) {
                  float x = __uint_as_float(r[i]);
                  if (p.bias) x += __ldg(p.bias + n + i);
                  if (res) x += DT<T>::to_f(res[(size_t)m * p.ldr + n + i]);
                  out[(size_t)m * p.ldo + n + i] = DT<T>::from_f(x);
                }
              }
            }
          }
        }
        release_acc(acc);
      }
    }
  } else if (warp == Cfg::STORE_WARP) {
    // =================================== store warp ===================================
    // Walks the output-producing work items of this CTA in order.  Item o uses staging buffer o & 1:
    //   [residual of item o TMA-loaded into it] -> epilogue warps turn it into the output tile -> stg_full[o & 1]
    //   -> TMA stores -> (reads done) -> residual of item o + 2 loaded into the same buffer, or stg_free[o & 1] raised.
    // The bulk async-groups belong to this thread, so only this thread ever waits for a store.
    if (p.epi_tma && RCDM_GEMM_EXPERIMENT != 5 && RCDM_GEMM_EXPERIMENT != 9 && elect_one()) {
      const bool has_res = p.res != nullptr && RCDM_GEMM_EXPERIMENT != 6;
      const int n_total = p.geglu ? p.N / 2 : p.N;
      const int wcols = p.geglu ? Cfg::QW / 2 : Cfg::QW;
      const int tile_cols = p.geglu ? BN / 2 : BN;
      constexpr int NG = Cfg::NG;
      auto issue_residual = [&](int t, int o) {  // residual tile of output item o (tile index t) -> staging[o & 1]
        const int nt = t % p.num_n_tiles;
        const int mt = PAIR ? 2 * (t / p.num_n_tiles) + (int)rank : t / p.num_n_tiles;
        const int b = o & 1;
        int parts = 0;
        for (int g = 0; g < NG; ++g) parts += (nt * BN + g * Cfg::QW) < p.N;
        mbar_expect_tx(&res_full[b], parts * Cfg::PART_BYTES);
        for (int g = 0; g < parts; ++g)
          tma_load_2d(staging + b * Cfg::STG_BYTES + g * Cfg::PART_BYTES, &maps.r, &res_full[b], nt * BN + g * Cfg::QW,
                      mt * 128);
      };
      GemmWork work(p, wid, nworkers);
      GemmWork peek(p, wid, nworkers);  // runs two output items ahead of `work`
      int pt, pk0, pk1;
      auto next_output = [&](GemmWork& w, int& t) {  // next work item that produces an output tile (kb0 == 0)
        int k0, k1;
        while (w.next(t, k0, k1))
          if (k0 == 0) return true;
        return false;
      };
      (void)pk0;
      (void)pk1;
      bool have_peek = next_output(peek, pt);
      if (has_res && have_peek) issue_residual(pt, 0);
      if (have_peek) have_peek = next_output(peek, pt);
      if (has_res && have_peek) issue_residual(pt, 1);
      int tile;
      for (int o = 0; next_output(work, tile); ++o) {
        const int n_tile = tile % p.num_n_tiles;
        const int m_tile = PAIR ? 2 * (tile / p.num_n_tiles) + (int)rank : tile / p.num_n_tiles;
        const int sb = o & 1;
        mbar_wait(&stg_full[sb], (o >> 1) & 1);
        GEMM_STAMP(o, 12);
#if RCDM_GEMM_EXPERIMENT != 6
        if (p.out4d) {  // conv tile origin on the output grid (same decomposition as the producer's)
          int t = m_tile;
          const int tx = t % p.tiles_x;
          t /= p.tiles_x;
          const int ty = t % p.tiles_y, tb = t / p.tiles_y;
          for (int g = 0; g < NG; ++g)
            if (n_tile * tile_cols + g * wcols < n_total)
              tma_store_4d(&maps.o, staging + sb * Cfg::STG_BYTES + g * Cfg::PART_BYTES, n_tile * tile_cols + g * wcols,
                           tx * p.tw, ty * p.th, tb * p.tn);
        } else {
          for (int g = 0; g < NG; ++g)
            if (n_tile * tile_cols + g * wcols < n_total)
              tma_store_2d(&maps.o, staging + sb * Cfg::STG_BYTES + g * Cfg::PART_BYTES, n_tile * tile_cols + g * wcols,
                           m_tile * 128);
        }
        bulk_commit();
        GEMM_STAMP(o, 13);
        bulk_wait_read0();  // the buffer may be refilled
        GEMM_STAMP(o, 14);
#endif
        if (have_peek) have_peek = next_output(peek, pt);  // item o + 2
        if (has_res) {
          if (have_peek) issue_residual(pt, o + 2);
        } else {
          mbar_arrive(&stg_free[sb]);
        }
        GEMM_STAMP(o, 15);
      }
      bulk_wait0();  // every TMA store of this CTA has completed before its shared memory goes away
    }
  }
  tc_fence_before();
  __syncwarp();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA may exit while its peer can still touch its smem / barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace rcdm
