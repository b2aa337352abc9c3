#!/usr/bin/env python
"""Benchmark of the RCDMs stage-2 denoise hot path on B200 (contract: see the task statement / DESIGN.md §5).

    python bench.py --gpus 1 --steps 3 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # the reference algorithm on the host CPU cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU, clips sharded (weak scaling)

A "step" is one pass of the hot path over one batch: the full denoise of `--clips` 5-frame clips per GPU
(BASELINE.json configs[1]: PororoSV 512x512 -> 64x64 latents, 50 DDIM steps, fp16, CFG guidance 2.0, L = 85)
from prepared latents / mask / masked-image latents / fused context to the final latents, by
`rcdm_denoise_loop` (CUDA-graph replay of UNet + CFG + DDIM).  metric = story-frames/sec = 5 * clips * N / t.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FORWARD_64 = 11.044e12  # algorithmic FLOPs of one UNet forward, 512^2, 1 clip, CFG, L=85
METRIC = "story-frames/sec @50 DDIM steps, 512x512x5-frame clip"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=1, help="clips per GPU (batched into one UNet call)")
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--latent", type=int, default=64, help="latent height = width (64 <=> 512x512 frames)")
    ap.add_argument("--ctx-len", type=int, default=85)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--guidance", type=float, default=2.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    ap.add_argument("--no-eager-gpu-baseline", action="store_true",
                    help="skip the eager-GPU leg (the oracle restatement in torch eager, bench dtype, on this GPU: the "
                         "stand-in for 'the reference modules in eager fp16 on the B200', SURVEY 8d; on by default at N=1)")
    ap.add_argument("--no-pixels", action="store_true", help="skip the end-to-end-to-pixels leg (denoise + VAE decode)")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the extra BASELINE configs (N=1: config 3 bf16 x 8 clips; N>1: config 5, 8 clips per GPU)")
    ap.add_argument("--eager-gpu-only", action="store_true", help="run only the eager-GPU baseline leg and exit")
    ap.add_argument("--workload", default="stage2", choices=["stage2", "prior"],
                    help="stage2 (default, the headline metric) | prior: BASELINE config 4, the stage-1 frame-prior loop "
                         "(SURVEY 8f rank 1), single GPU")
    ap.add_argument("--lib-option", action="append", default=[], metavar="NAME=VALUE",
                    help="A/B legs only: library-wide debug switch (rcdm_debug_set_option), e.g. attn_short_kv=0")
    ap.add_argument("--unet-option", action="append", default=[], metavar="NAME=VALUE",
                    help="A/B legs only: per-handle debug switch of the UNet (rcdm_unet_set_option), e.g. po_fold=0")
    ap.add_argument("--prior-steps", type=int, default=100)
    ap.add_argument("--prior-two-gemm-proj-out", action="store_true",
                    help="prior workload: ff.net.2 and proj_out of the motion modules as two GEMMs (A/B leg)")
    ap.add_argument("--prior-standalone-ln", action="store_true",
                    help="prior workload: every nn.LayerNorm as its own launch instead of folded around the GEMMs (A/B leg)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1400.0), tflops_burst=d.get("bf16_tflops", 1590.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_step_seconds(latent, ctx_len, ddim_steps, guidance, n_steps, n_warm, budget_s):
    """Times `n_steps` DDIM steps (UNet forward at `latent`^2 + CFG + scheduler step) of ONE clip with the
    oracle restatement of the reference (fp32, all host threads).  Returns (mean seconds per DDIM step, info)."""
    import torch
    from oracle.loop_ref import make_scheduler
    from oracle.unet_ref import unet_forward
    from rcdms_b200.synthetic import synthetic_clip_inputs, synthetic_state_dict
    from rcdms_b200.unet_spec import full_config
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = full_config()
    sd = synthetic_state_dict(cfg, seed=0)
    inp = synthetic_clip_inputs(0, latent, latent, ctx_len)
    sched = make_scheduler()
    sched.set_timesteps(ddim_steps)
    lat, mask, ml, ctx = inp["latents"], inp["mask"], inp["masked_latents"], inp["ctx"]
    mask2, ml2 = torch.cat([mask] * 2), torch.cat([ml] * 2)
    times = []
    t_start = time.time()
    with torch.no_grad():
        for i, t in enumerate(sched.timesteps[: n_warm + n_steps]):
            t0 = time.time()
            x = torch.cat([torch.cat([lat] * 2), mask2, ml2], dim=1)
            eps = unet_forward(sd, cfg, x, t, ctx)
            eu, ec = eps.chunk(2)
            lat = sched.step(eu + guidance * (ec - eu), t, lat, eta=0.0).prev_sample
            dt = time.time() - t0
            if i >= n_warm:
                times.append(dt)
            if time.time() - t_start > budget_s and times:
                break
    return sum(times) / len(times), dict(cores=cores, threads=torch.get_num_threads(), executed=len(times))


def run_reference(a):
    """Reference arm: the reference's algorithm (oracle port; the Python reference cannot travel to the GPU box) on the
    host CPU cores, same workload / metric / unit.  One "step" = a BOUNDED SAMPLE of one clip's denoise: ONE of its
    `ddim_steps` DDIM steps (fp32 UNet forward with CFG at the configured latent size + scheduler step), ~10 s on 16
    cores; value extrapolates the measured sample to the whole clip (x ddim_steps).  ms_per_step is the MEASURED duration
    of a sample, so steps x ms_per_step is the wall time of the timed region."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    t_step, info = cpu_reference_step_seconds(a.latent, a.ctx_len, a.ddim_steps, a.guidance, a.steps,
                                              min(a.warmup, 1), budget_s=1e9)
    clip_s = t_step * a.ddim_steps
    value = 5.0 / clip_s
    sample = (f"each step = 1 of the {a.ddim_steps} DDIM steps of one clip (UNet fp32 forward at {a.latent}x{a.latent} "
              f"latents with CFG + DDIM step), {info['executed']} steps timed, {t_step:.2f} s each; value = 5 frames / "
              f"({a.ddim_steps} x that)")
    line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=a.gpus, steps=info["executed"],
                warmup=min(a.warmup, 1), ms_per_step=t_step * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference", config=config_dict(a, world),
                sample_fraction_of_step_unit=1.0 / a.ddim_steps,
                note="reference algorithm = oracle port of src/models/unet.py + diffusers DDIM on the host CPU cores; "
                     "a step here is a bounded sample (1 DDIM step) of the clip the GPU arm denoises per step",
                cpu_baseline=dict(value=value, unit="frames/s", cores=info["cores"], kind="port", sample=sample),
                e2e=dict(value=value, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def workload_name(a, clips=None, dtype=None, ctx_len=None):
    px = a.latent * 8
    ctx_len = ctx_len or a.ctx_len
    ds = "FlintstonesSV" if ctx_len == 91 else "PororoSV"
    return (f"stage2 {ds} {px}x{px}, {a.ddim_steps} DDIM steps, {dtype or a.dtype}, batch={clips or a.clips} clip(s)/GPU, "
            f"CFG {a.guidance}, L={ctx_len}")


def config_dict(a, world):
    """The workload description - identical keys and values in both arms (ours / --impl reference) for the same flags."""
    return dict(workload=workload_name(a), clips_per_gpu=a.clips, ddim_steps=a.ddim_steps, latent=a.latent,
                ctx_len=a.ctx_len, guidance=a.guidance,
                parallelism=f"clip-sharded x{world}, one all_gather of final latents",
                l2="not flushed: per-step working set (2.55 GB fp16 weights + activations) >> 126 MB L2")


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def eager_gpu_leg(a, dtype, n_e=3):
    """"the reference modules in PyTorch eager fp16 on the B200" - the denominator of the north_star's
    10x target.  The Python reference cannot travel to the GPU box, so its restatement (oracle/unet_ref.py: the same
    torch op sequence, one cuBLAS / cuDNN launch per op, attention scores materialised; pinned to the reference modules
    by tests/golden) runs in the bench dtype on this GPU.  A reported baseline like cpu_baseline, never a product path."""
    import torch
    from oracle.loop_ref import make_scheduler
    from oracle.unet_ref import unet_forward
    from rcdms_b200.synthetic import synthetic_clip_inputs, synthetic_state_dict
    from rcdms_b200.unet_spec import full_config
    ecfg = full_config()
    esd = {k: v.to("cuda", dtype) for k, v in synthetic_state_dict(ecfg, seed=0).items()}
    ein = {k: v.to("cuda", dtype) for k, v in synthetic_clip_inputs(0, a.latent, a.latent, a.ctx_len).items()}
    sch = make_scheduler()
    sch.set_timesteps(a.ddim_steps)
    lat = ein["latents"]
    m2, l2 = torch.cat([ein["mask"]] * 2), torch.cat([ein["masked_latents"]] * 2)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for i, t in enumerate(sch.timesteps[: n_e + 1]):
            if i == 1:
                ev0.record()
            x = torch.cat([torch.cat([lat] * 2), m2, l2], dim=1)
            eps = unet_forward(esd, ecfg, x, t, ein["ctx"])
            eu, ec = eps.chunk(2)
            lat = sch.step(eu + a.guidance * (ec - eu), t, lat, eta=0.0).prev_sample
        ev1.record()
    torch.cuda.synchronize()
    ms_e = ev0.elapsed_time(ev1) / n_e
    del esd
    torch.cuda.empty_cache()
    return dict(value=5.0 / (ms_e * a.ddim_steps / 1e3), unit="frames/s", ms_per_ddim_step=ms_e, kind="port",
                sample=f"{n_e} of {a.ddim_steps} DDIM steps of one clip after 1 warm-up step, oracle restatement of the "
                       f"reference UNet + CFG + DDIM in torch eager {a.dtype} on this GPU (cuBLAS / cuDNN, one launch per "
                       f"op), extrapolated x{a.ddim_steps}; the reference's own nn.Modules are Python and absent on this box")


def run_ours(a):
    import torch
    import torch.distributed as dist
    from rcdms_b200 import _lib
    from rcdms_b200.models import UNet3DConditionModel
    from rcdms_b200.parallel import gather_latents
    from rcdms_b200.pipelines.RCDMs_pipeline import RCDMsPipeline
    from rcdms_b200.schedulers import DDIMScheduler
    from rcdms_b200.synthetic import synthetic_clip_inputs, synthetic_state_dict
    from rcdms_b200.unet_spec import RCDMS_SCHEDULER_KWARGS, full_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()  # fails loudly if the CUDA library is missing
    for kv in a.lib_option:
        if L.rcdm_debug_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1])) < 0:
            raise SystemExit(f"unknown library option {kv}")
    cfg = full_config()
    sd = synthetic_state_dict(cfg, seed=0)

    class _VaeCfg:  # the pipeline only needs vae_scale_factor = 8 here; VAE/CLIP stay outside the hot path
        block_out_channels = (128, 256, 512, 512)

    class _Vae:
        config = _VaeCfg()

    def make_pipe(dtype):
        unet = UNet3DConditionModel.from_config(cfg)
        for kv in a.unet_option:
            unet.set_debug_option(kv.split("=")[0], int(kv.split("=")[1]))
        unet.load_state_dict(sd, strict=True)
        unet = unet.to(device=dev, dtype=dtype)
        pipe = RCDMsPipeline(vae=_Vae(), text_encoder=None, tokenizer=None, unet=unet, local_module=None,
                             global_module=None, scheduler=DDIMScheduler(**RCDMS_SCHEDULER_KWARGS))
        pipe.use_cuda_graph = not a.no_graph
        return unet, pipe

    def measure(pipe, dtype, clips, ctx_len, steps, warm):
        """(ms resident, ms e2e, host input bytes, host output bytes, launches) for `steps` denoises of `clips` clips per GPU."""
        # per-clip synthetic inputs, clip index = global (rank-independent results); pinned host copies for e2e
        clip_ids = [rank * clips + i for i in range(clips)]
        ins = [synthetic_clip_inputs(k, a.latent, a.latent, ctx_len) for k in clip_ids]
        host = dict(
            latents=torch.cat([i["latents"] for i in ins]).to(dtype).pin_memory(),
            masked=torch.cat([i["masked_latents"] for i in ins]).to(dtype).pin_memory(),
            mask=torch.cat([i["mask"] for i in ins]).to(dtype).pin_memory(),
            # ctx rows ordered (b f) with b = [uncond clips..., cond clips...]
            ctx=torch.cat([i["ctx"][:5] for i in ins] + [i["ctx"][5:] for i in ins]).to(dtype).pin_memory())
        devt = {k: v.to(dev) for k, v in host.items()}
        out_host = torch.empty_like(host["latents"]).pin_memory()

        def denoise_resident():
            return pipe.denoise(devt["latents"], torch.cat([devt["mask"]] * 2), torch.cat([devt["masked"]] * 2),
                                devt["ctx"], a.ddim_steps, a.guidance)

        def denoise_e2e():
            d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            lat = pipe.denoise(d["latents"], torch.cat([d["mask"]] * 2), torch.cat([d["masked"]] * 2), d["ctx"],
                               a.ddim_steps, a.guidance)
            out_host.copy_(lat, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return lat

        def gather(lat):  # the path's only collective: final latents of every shard
            return gather_latents(lat, clips * world)

        def timed(fn, k):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                gather(fn())
            e1.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return ms.item()

        for _ in range(warm):
            gather(denoise_resident())
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = L.rcdm_kernel_launches()
        ms_res = timed(denoise_resident, steps)
        launches = L.rcdm_kernel_launches() - l0
        clocks = sampler.stop() if rank == 0 else None
        gather(denoise_e2e())
        ms_e2e = timed(denoise_e2e, steps)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = out_host.numel() * out_host.element_size()
        return dict(ms=ms_res / steps, ms_e2e=ms_e2e / steps, h2d=h2d, d2h=d2h, launches=int(launches), clocks=clocks,
                    devt=devt)

    dtype = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    unet, pipe = make_pipe(dtype)
    warm = max(a.warmup, 3)
    r = measure(pipe, dtype, a.clips, a.ctx_len, a.steps, warm)
    devt = r["devt"]
    frames = 5 * a.clips * world
    value = frames / (r["ms"] / 1e3)
    e2e_value = frames / (r["ms_e2e"] / 1e3)
    flop_step = FLOP_PER_FORWARD_64 * (a.latent / 64.0) ** 2 * a.clips * a.ddim_steps  # per GPU per "step" (approx. off 64^2)
    pk = peaks()

    roofline = None
    if not a.no_profile and rank == 0:
        # dominant kernel = gemm_tcgen05_kernel (Linear / conv1x1 / implicit-GEMM conv3x3): per-launch CUDA-event timing
        x = torch.cat([torch.cat([devt["latents"]] * 2), torch.cat([devt["mask"]] * 2), torch.cat([devt["masked"]] * 2)], 1)
        prof = unet.profile(x, 501.0, devt["ctx"], reps=3)
        agg = {}
        for p in prof:
            g = agg.setdefault(p["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
            g["ms"] += p["ms"]
            g["flops"] += p["flops"]
            g["bytes"] += p["bytes"]
            g["n"] += 1
        tot_ms = sum(g["ms"] for g in agg.values())
        mm = [agg[k] for k in ("gemm", "gemm_geglu", "conv3x3") if k in agg]
        mm_ms, mm_fl, mm_n = sum(g["ms"] for g in mm), sum(g["flops"] for g in mm), sum(g["n"] for g in mm)
        achieved = mm_fl / (mm_ms / 1e3) / 1e12
        traffic, traffic_note = None, None
        for tname in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
            tp = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu pass
                tj = json.load(open(tp))
                traffic, traffic_note = tj.get("dram_bytes_per_launch"), f"profiles/{tname}: " + str(tj.get("note"))
                break
        roofline = dict(bound="tensor", kernel="gemm_tcgen05_kernel", achieved=achieved, peak=pk["tflops"], unit="TFLOP/s",
                        frac=achieved / pk["tflops"], traffic=traffic, traffic_note=traffic_note,
                        peak_source=pk["source"] + ", sustained bf16",
                        launches_per_forward=mm_n, share_of_forward=mm_ms / tot_ms,
                        flops_per_launch_avg=mm_fl / mm_n, ms_per_launch_avg=mm_ms / mm_n,
                        by_kind={k: dict(ms=round(v["ms"], 4), n=v["n"],
                                         tflops=round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1),
                                         gbs=round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)) for k, v in agg.items()},
                        forward_ms_sum_of_ops=tot_ms)

    # ---- end to end to PIXELS (SURVEY 8d "also report end-to-end including VAE decode"): pinned host inputs -> denoise ->
    # AutoencoderKL.decode of all frames on the same kernels (rcdms_b200.models.vae) -> clamp -> fp32 frames on the host
    e2e_pixels = None
    if not a.no_pixels and rank == 0 and world == 1 and a.latent == 64:
        from rcdms_b200.models import AutoencoderKL
        from rcdms_b200.vae_spec import synthetic_vae_state_dict, vae_full_config
        vcfg = vae_full_config()
        vae = AutoencoderKL.from_config(vcfg)
        vae.load_state_dict(synthetic_vae_state_dict(vcfg, seed=0), strict=True)
        vae = vae.to(device=dev, dtype=dtype)
        ins = [synthetic_clip_inputs(k, a.latent, a.latent, a.ctx_len) for k in range(a.clips)]
        hostp = dict(latents=torch.cat([i["latents"] for i in ins]).to(dtype).pin_memory(),
                     masked=torch.cat([i["masked_latents"] for i in ins]).to(dtype).pin_memory(),
                     mask=torch.cat([i["mask"] for i in ins]).to(dtype).pin_memory(),
                     ctx=torch.cat([i["ctx"][:5] for i in ins] + [i["ctx"][5:] for i in ins]).to(dtype).pin_memory())
        px = a.latent * 8
        pix_host = torch.empty((a.clips * 5, 3, px, px), dtype=torch.float32).pin_memory()

        def to_pixels():
            d = {k: v.to(dev, non_blocking=True) for k, v in hostp.items()}
            lat = pipe.denoise(d["latents"], torch.cat([d["mask"]] * 2), torch.cat([d["masked"]] * 2), d["ctx"],
                               a.ddim_steps, a.guidance)
            frames = (lat / 0.18215).permute(0, 2, 1, 3, 4).flatten(0, 1)          # RCDMs_pipeline.py:276-278
            video = (vae.decode(frames).sample / 2 + 0.5).clamp(0, 1)               # :281-284 (all frames in one batch)
            pix_host.copy_(video.float(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return frames

        frames = to_pixels()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            to_pixels()
        e1.record()
        torch.cuda.synchronize()
        ms_pix = e0.elapsed_time(e1) / a.steps
        e0.record()
        for _ in range(3):
            vae.decode(frames)
        e1.record()
        torch.cuda.synchronize()
        e2e_pixels = dict(value=5 * a.clips / (ms_pix / 1e3), unit="frames/s", ms_per_step=ms_pix,
                          vae_decode_ms=e0.elapsed_time(e1) / 3,
                          h2d_bytes_per_step=sum(v.numel() * v.element_size() for v in hostp.values()),
                          d2h_bytes_per_step=pix_host.numel() * pix_host.element_size(),
                          api="RCDMsPipeline.denoise + AutoencoderKL.decode (rcdms_b200 kernels) + clamp; host pinned "
                              "tensors in, fp32 frames (clips*5, 3, 512, 512) on the host out",
                          vae="SD-1.5 AutoencoderKL architecture (83.65 M parameters), synthetic weights")
        del vae, frames
        torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations, measured in the same run (reported, not the headline):
    #   N = 1: config 3 (FlintstonesSV, bf16, 8 clips batched on the GPU, L = 91);  N > 1: config 5 (8 clips per GPU, fp16)
    extra = []
    if not a.no_extra_configs and a.clips == 1 and a.latent == 64:
        del devt, r["devt"]
        if world == 1:
            del unet, pipe
            torch.cuda.empty_cache()
            _, pipe3 = make_pipe(torch.bfloat16)
            e = measure(pipe3, torch.bfloat16, 8, 91, 1, 2)
            extra.append(dict(workload=workload_name(a, 8, "bf16", 91), baseline_config=3, dtype="bf16",
                              value=5 * 8 / (e["ms"] / 1e3), unit="frames/s", ms_per_step=e["ms"],
                              ms_per_ddim_step=e["ms"] / a.ddim_steps, e2e_value=5 * 8 / (e["ms_e2e"] / 1e3),
                              clocks=e["clocks"], steps=1, warmup=2))
            del pipe3
            torch.cuda.empty_cache()
            extra.append(prior_extra_config(a))
        else:
            e = measure(pipe, dtype, 8, a.ctx_len, 1, 2)
            extra.append(dict(workload=workload_name(a, 8) + f", {8 * world} clips over {world} GPUs", baseline_config=5,
                              dtype=a.dtype, value=5 * 8 * world / (e["ms"] / 1e3), unit="frames/s", ms_per_step=e["ms"],
                              ms_per_ddim_step=e["ms"] / a.ddim_steps, e2e_value=5 * 8 * world / (e["ms_e2e"] / 1e3),
                              clocks=e["clocks"], steps=1, warmup=2))

    cpu_baseline = None
    if not a.no_cpu_baseline and rank == 0 and world == 1:
        t_step, info = cpu_reference_step_seconds(a.latent, a.ctx_len, a.ddim_steps, a.guidance, 1, 0, a.cpu_budget_s)
        cpu_baseline = dict(value=5.0 / (t_step * a.ddim_steps), unit="frames/s", cores=info["cores"], kind="port",
                            sample=f"1 of {a.ddim_steps} DDIM steps (UNet fp32 forward at {a.latent}x{a.latent} latents + "
                                   f"CFG + DDIM) of one clip, {t_step:.1f} s, extrapolated x{a.ddim_steps}")

    eager_gpu = None
    if not a.no_eager_gpu_baseline and rank == 0 and world == 1:
        eager_gpu = eager_gpu_leg(a, dtype)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=a.steps, warmup=warm,
                    ms_per_step=r["ms"], higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f16" if a.dtype == "fp16" else "bf16", data="synthetic", config=config_dict(a, world),
                    derived=dict(cuda_graph=not a.no_graph, ms_per_ddim_step=r["ms"] / a.ddim_steps,
                                 achieved_tflops_per_gpu=flop_step / (r["ms"] / 1e3) / 1e12,
                                 frac_of_sustained_bf16_peak=flop_step / (r["ms"] / 1e3) / 1e12 / pk["tflops"]),
                    e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=r["h2d"], d2h_bytes_per_step=r["d2h"],
                             ms_per_step=r["ms_e2e"], api="RCDMsPipeline.denoise (host pinned tensors in, host latents out)"),
                    gpu_launches=r["launches"], clocks=r["clocks"], roofline=roofline, cpu_baseline=cpu_baseline)
        if eager_gpu is not None:
            line["gpu_eager_baseline"] = eager_gpu
            line["derived"]["speedup_vs_gpu_eager"] = eager_gpu["ms_per_ddim_step"] / (r["ms"] / a.ddim_steps)
        if e2e_pixels is not None:
            line["e2e_pixels"] = e2e_pixels
        if extra:
            line["extra_configs"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------
# --workload prior: BASELINE config 4 (stage-1 frame-prior transformer diffusion, 100 UnCLIP steps, 1 GPU)
# ----------------------------------------------------------------------------------------------------------
PRIOR_METRIC = "stage-1 prior frame-embeddings/sec @100 UnCLIP steps, 5-frame clip"


def prior_cpu_seconds_per_forward(layers_sample=2):
    """Oracle restatement of MyPriorTransformer.forward (fp32, all host threads) on a `layers_sample`-layer slice of
    the shipped configuration, scaled to the 20 layers (the layers are identical in cost)."""
    import torch
    from oracle.prior_ref import prior_forward
    from rcdms_b200.prior_spec import prior_full_config, prior_state_dict_spec
    from rcdms_b200.synthetic import synthetic_prior_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = prior_full_config(num_layers=layers_sample)
    g = torch.Generator().manual_seed(0)
    sd = {n: torch.randn(sh, generator=g) * 0.02 for n, sh in prior_state_dict_spec(cfg)}
    inp = synthetic_prior_inputs(cfg, 0)
    args = (torch.cat([inp["latents"]] * 2), 500, inp["prompt_embeds"], inp["text_hidden"],
            torch.cat([inp["imgs_proj_embeds1"]] * 2), torch.cat([inp["mask_label"]] * 2), inp["text_mask"])
    with torch.no_grad():
        prior_forward(sd, cfg, *args)
        t0 = time.time()
        prior_forward(sd, cfg, *args)
        dt = time.time() - t0
    full = prior_full_config()["num_layers"]
    return dt * full / layers_sample, dict(cores=cores, sample_s=dt, layers_sample=layers_sample)


def prior_extra_config(a, steps=1, warm=2):
    """BASELINE config 4 (stage-1 frame prior, 100 UnCLIP steps, one clip) measured inside the default stage-2 run so that the
    driver's record carries it: device-resident value only (the full line with e2e / cpu_baseline: --workload prior)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from bench_prior import algorithmic_flops, device_random_weights
    from rcdms_b200.models.myprior_transformer import MyPriorTransformer
    from rcdms_b200.pipelines.prior_pipeline import Seq_Inpaint_Prior_Pipeline
    from rcdms_b200.prior_spec import PRIOR_SCHEDULER_KWARGS, prior_full_config
    from rcdms_b200.schedulers import UnCLIPScheduler
    from rcdms_b200.synthetic import stack_prior_clips, synthetic_prior_inputs
    dtype = torch.float16
    steps_n = a.prior_steps
    cfg = prior_full_config()
    model = device_random_weights(MyPriorTransformer.from_config(cfg), dtype)
    pipe = Seq_Inpaint_Prior_Pipeline(prior=model, image_encoder=None, text_encoder=None, tokenizer=None,
                                      scheduler=UnCLIPScheduler(**PRIOR_SCHEDULER_KWARGS))
    pipe.use_cuda_graph = not a.no_graph
    dev = {k: (v.to(dtype) if v.is_floating_point() else v).cuda()
           for k, v in stack_prior_clips([synthetic_prior_inputs(cfg, 0)]).items()}
    gen = torch.Generator(device="cuda").manual_seed(42)

    def sample():
        return pipe.sample(dev["latents"], dev["prompt_embeds"], dev["text_hidden"], dev["text_mask"],
                           dev["imgs_proj_embeds1"], dev["mask_label"], steps_n, 4.0, generator=gen)

    for _ in range(warm):
        sample()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sample()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.stop()
    tf = algorithmic_flops(cfg, 10) * steps_n / (ms / 1e3) / 1e12
    del pipe, model
    torch.cuda.empty_cache()
    return dict(workload="stage-1 prior: kandinsky-2-2 prior + 20 prior-state motion modules (2.88 B params), 97 tokens, CFG 4.0, "
                         f"{steps_n} UnCLIP steps, 1 clip per run", baseline_config=4, dtype="f16", value=5.0 / (ms / 1e3),
                unit="frame-embeddings/s", ms_per_step=ms, ms_per_unclip_step=ms / steps_n, achieved_tflops=tf, clocks=clocks,
                steps=steps, warmup=warm)


def run_prior(a):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from bench_prior import algorithmic_flops, device_random_weights
    from rcdms_b200 import _lib
    from rcdms_b200.models.myprior_transformer import MyPriorTransformer
    from rcdms_b200.pipelines.prior_pipeline import Seq_Inpaint_Prior_Pipeline
    from rcdms_b200.prior_spec import PRIOR_SCHEDULER_KWARGS, prior_full_config
    from rcdms_b200.schedulers import UnCLIPScheduler
    from rcdms_b200.synthetic import synthetic_prior_inputs
    steps_n = a.prior_steps
    if a.impl == "reference":
        t_fwd, info = prior_cpu_seconds_per_forward()
        v = 5.0 / (t_fwd * steps_n)
        cb = dict(value=v, unit="frame-embeddings/s", cores=info["cores"], kind="port",
                  sample=f"oracle forward of a {info['layers_sample']}-layer slice ({info['sample_s']:.1f} s), scaled to 20 "
                         f"layers x {steps_n} steps")
        print(json.dumps(dict(metric=PRIOR_METRIC, value=v, unit="frame-embeddings/s", n_gpus=1, steps=1, warmup=1,
                              ms_per_step=t_fwd * steps_n * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                              dtype="f32", data="synthetic", impl="reference",
                              config=dict(workload="stage-1 prior, kandinsky-2-2 + motion modules, CFG 4.0"),
                              cpu_baseline=cb, gpu_launches=0,
                              e2e=dict(value=v, unit="frame-embeddings/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return
    pk = peaks()
    dtype = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    cfg = prior_full_config()
    model = device_random_weights(MyPriorTransformer.from_config(cfg), dtype)
    model.fold_layernorm = not a.prior_standalone_ln
    model.fold_proj_out = not a.prior_two_gemm_proj_out
    pipe = Seq_Inpaint_Prior_Pipeline(prior=model, image_encoder=None, text_encoder=None, tokenizer=None,
                                      scheduler=UnCLIPScheduler(**PRIOR_SCHEDULER_KWARGS))
    pipe.use_cuda_graph = not a.no_graph
    from rcdms_b200.synthetic import stack_prior_clips
    host = {k: (v.to(dtype) if v.is_floating_point() else v).pin_memory()
            for k, v in stack_prior_clips([synthetic_prior_inputs(cfg, i) for i in range(a.clips)]).items()}
    gen = torch.Generator(device="cuda").manual_seed(42)

    def sample(dev):
        return pipe.sample(dev["latents"], dev["prompt_embeds"], dev["text_hidden"], dev["text_mask"],
                           dev["imgs_proj_embeds1"], dev["mask_label"], steps_n, 4.0, generator=gen)

    dev = {k: v.cuda() for k, v in host.items()}
    warm = max(a.warmup, 3)
    for _ in range(warm):
        sample(dev)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = sample(dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches_total = pipe.last_gpu_launches * a.steps  # graph replays included (counted by the pipeline)
    out_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    e0.record()
    for _ in range(a.steps):
        d = {k: v.cuda(non_blocking=True) for k, v in host.items()}
        out_host.copy_(model.post_process_latents(sample(d)), non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / a.steps
    clocks = sampler.stop()
    fl = algorithmic_flops(cfg, 10 * a.clips) * steps_n
    frames = 5.0 * a.clips
    cb = None
    if not a.no_cpu_baseline:
        t_fwd, info = prior_cpu_seconds_per_forward()
        cb = dict(value=5.0 / (t_fwd * steps_n), unit="frame-embeddings/s", cores=info["cores"], kind="port",
                  sample=f"oracle forward of a {info['layers_sample']}-layer slice ({info['sample_s']:.1f} s), scaled to 20 "
                         f"layers x {steps_n} steps (one clip)")
    tf = fl / (ms / 1e3) / 1e12
    print(json.dumps(dict(
        metric=PRIOR_METRIC, value=frames / (ms / 1e3), unit="frame-embeddings/s", n_gpus=1, steps=a.steps, warmup=warm,
        ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f16" if a.dtype == "fp16" else "bf16", data="synthetic",
        config=dict(workload="stage-1 prior: kandinsky-2-2 prior + 20 prior-state motion modules (2.88 B params), 97 tokens, "
                             f"CFG 4.0 (10 rows per clip), {steps_n} UnCLIP steps, {a.clips} clip(s) per run", clips_per_gpu=a.clips,
                    cuda_graph=not a.no_graph, layernorm="standalone launches" if a.prior_standalone_ln else "folded around the GEMMs",
                    proj_out="two GEMMs" if (a.prior_two_gemm_proj_out or a.prior_standalone_ln) else "folded over ff.net.2",
                    launches_per_unclip_step=int(pipe.last_gpu_launches // steps_n),
                    ms_per_unclip_step=ms / steps_n,
                    l2="not flushed: 5.76 GB of fp16 weights stream through per step >> 126 MB L2"),
        e2e=dict(value=frames / (ms_e2e / 1e3), unit="frame-embeddings/s", ms_per_step=ms_e2e,
                 h2d_bytes_per_step=sum(v.numel() * v.element_size() for v in host.values()),
                 d2h_bytes_per_step=out_host.numel() * out_host.element_size(),
                 api="Seq_Inpaint_Prior_Pipeline.sample + post_process_latents (host pinned tensors in, host embeddings out)"),
        gpu_launches=int(launches_total), clocks=clocks,
        roofline=dict(bound="tensor", kernel="gemm_tcgen05_kernel", achieved=tf, peak=pk["tflops"], unit="TFLOP/s",
                      frac=tf / pk["tflops"], traffic=None, peak_source=pk["source"],
                      note="achieved = algorithmic FLOPs of the whole step / whole-step time (lower bound for the GEMM "
                           "kernel, which is ~85 % of the step: profiles/r01_prior_*)"),
        cpu_baseline=cb)), flush=True)


def run_eager_gpu_only_prior(a):
    """--eager-gpu-only --workload prior: oracle restatement of the prior loop in torch eager on this GPU."""
    import torch
    from oracle.prior_ref import make_prior_scheduler, prior_forward
    from rcdms_b200.prior_spec import prior_full_config, prior_state_dict_spec
    from rcdms_b200.synthetic import positional_encoding, synthetic_prior_inputs
    dtype = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    cfg = prior_full_config()
    g = torch.Generator(device="cuda").manual_seed(0)
    sd = {}
    for name, shape in prior_state_dict_spec(cfg):
        if name.endswith("pos_encoder.pe"):
            sd[name] = positional_encoding(shape[1], shape[2]).to("cuda", dtype)
            continue
        is_norm = any(k in "." + name for k in (".norm", "norms.", "ff_norm", "prior_norm"))
        if is_norm:
            sd[name] = (torch.ones(shape, device="cuda") if name.endswith("weight") else torch.zeros(shape, device="cuda")).to(dtype)
            continue
        fan_in = shape[1] if (len(shape) == 2 and name.endswith("weight")) else 256
        sd[name] = ((torch.rand(shape, generator=g, device="cuda") * 2 - 1) * fan_in ** -0.5).to(dtype)
    inp = {k: (v.to("cuda", dtype) if v.is_floating_point() else v.cuda()) for k, v in synthetic_prior_inputs(cfg, 0).items()}
    sch = make_prior_scheduler()
    sch.set_timesteps(a.prior_steps)
    ts = sch.timesteps
    lat = inp["latents"]
    p1, ml = torch.cat([inp["imgs_proj_embeds1"]] * 2), torch.cat([inp["mask_label"]] * 2)
    gen = torch.Generator(device="cuda").manual_seed(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e = max(1, a.steps)
    with torch.no_grad():
        for i in range(n_e + 1):
            if i == 1:
                ev0.record()
            pred = prior_forward(sd, cfg, torch.cat([lat] * 2), ts[i], inp["prompt_embeds"], inp["text_hidden"], p1, ml,
                                 inp["text_mask"])
            pu, pt = pred.chunk(2)
            lat = sch.step(pu + 4.0 * (pt - pu), timestep=ts[i], sample=lat, generator=gen, prev_timestep=ts[i + 1]).prev_sample
        ev1.record()
    torch.cuda.synchronize()
    ms_e = ev0.elapsed_time(ev1) / n_e
    print(json.dumps(dict(metric=PRIOR_METRIC, gpu_eager_baseline=dict(
        value=5.0 / (ms_e * a.prior_steps / 1e3), unit="frame-embeddings/s", ms_per_unclip_step=ms_e, kind="port",
        dtype=a.dtype, finite=bool(torch.isfinite(lat).all()),
        sample=f"{n_e} of {a.prior_steps} UnCLIP steps of one clip after 1 warm-up step; oracle restatement of "
               f"MyPriorTransformer + CFG + UnCLIP in torch eager on this GPU (random weights), extrapolated"))), flush=True)


def run_eager_gpu_only(a):
    """--eager-gpu-only: just the eager-GPU baseline leg (random weights drawn on the device: timing only)."""
    import torch
    from oracle.loop_ref import make_scheduler
    from oracle.unet_ref import unet_forward
    from rcdms_b200.synthetic import synthetic_clip_inputs
    from rcdms_b200.unet_spec import BUFFER_SUFFIX, full_config, state_dict_spec
    dtype = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    cfg = full_config()
    g = torch.Generator(device="cuda").manual_seed(0)
    sd = {}
    for name, shape in state_dict_spec(cfg):
        if name.endswith(BUFFER_SUFFIX):
            from rcdms_b200.synthetic import positional_encoding
            sd[name] = positional_encoding(shape[1], shape[2]).to("cuda", dtype)
            continue
        fan_in = 1
        for d_ in shape[1:]:
            fan_in *= d_
        w = torch.randn(shape, generator=g, device="cuda") * (max(fan_in, 1) ** -0.5 if name.endswith("weight") and
                                                             len(shape) > 1 else 0.02)
        if len(shape) == 1 and name.endswith("weight"):
            w = 1 + w
        sd[name] = w.to(dtype)
    ein = {k: v.to("cuda", dtype) for k, v in synthetic_clip_inputs(0, a.latent, a.latent, a.ctx_len).items()}
    sch = make_scheduler()
    sch.set_timesteps(a.ddim_steps)
    lat = ein["latents"]
    m2, l2 = torch.cat([ein["mask"]] * 2), torch.cat([ein["masked_latents"]] * 2)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e = max(1, a.steps)
    with torch.no_grad():
        for i, t in enumerate(sch.timesteps[: n_e + 1]):
            if i == 1:
                ev0.record()
            x = torch.cat([torch.cat([lat] * 2), m2, l2], dim=1)
            eps = unet_forward(sd, cfg, x, t, ein["ctx"])
            eu, ec = eps.chunk(2)
            lat = sch.step(eu + a.guidance * (ec - eu), t, lat, eta=0.0).prev_sample
        ev1.record()
    torch.cuda.synchronize()
    ms_e = ev0.elapsed_time(ev1) / n_e
    print(json.dumps(dict(metric=METRIC, gpu_eager_baseline=dict(
        value=5.0 / (ms_e * a.ddim_steps / 1e3), unit="frames/s", ms_per_ddim_step=ms_e, kind="port", dtype=a.dtype,
        finite=bool(torch.isfinite(lat).all()),
        sample=f"{n_e} of {a.ddim_steps} DDIM steps of one clip after 1 warm-up step; oracle restatement of the reference "
               f"UNet + CFG + DDIM in torch eager on this GPU (random weights), extrapolated x{a.ddim_steps}"))), flush=True)


def main():
    a = parse()
    if a.eager_gpu_only:
        (run_eager_gpu_only_prior if a.workload == "prior" else run_eager_gpu_only)(a)
    elif a.workload == "prior":
        run_prior(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
