"""Stage-1 frame-prior sampling loop at the shipped size (BASELINE config 4):
kandinsky-2-2 prior + 20 prior-state motion modules (2.88 B parameters, inner width 2048, 97 tokens, CFG => 10 rows),
N UnCLIP steps through ``Seq_Inpaint_Prior_Pipeline.sample`` (CUDA graph of one step, replayed).

    python scripts/bench_prior.py [--steps 100] [--reps 3] [--layers 20] [--no-graph] [--once]

Prints one JSON line: frame-embeddings/s, ms per step, algorithmic TFLOP/s (2*MAC of every Linear + attention
matmuls, no padding).  Weights are random (drawn on the device: no checkpoint exists offline) — throughput only;
numerics are covered by tests/test_prior_gpu.py.  --once: a single un-graphed step (for ncu launch lists).
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200 import _lib  # noqa: E402
from rcdms_b200.models.myprior_transformer import MyPriorTransformer  # noqa: E402
from rcdms_b200.pipelines.prior_pipeline import Seq_Inpaint_Prior_Pipeline  # noqa: E402
from rcdms_b200.prior_spec import PRIOR_SCHEDULER_KWARGS, prior_dims, prior_full_config  # noqa: E402
from rcdms_b200.schedulers import UnCLIPScheduler  # noqa: E402
from rcdms_b200.synthetic import synthetic_prior_inputs  # noqa: E402


def device_random_weights(model, dtype, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    model.to(device="cuda", dtype=dtype)
    with torch.no_grad():
        for name, p in model.state_dict(keep_vars=True).items():
            if name.endswith("pos_encoder.pe"):
                continue
            is_norm = any(k in "." + name for k in (".norm", "norms.", "ff_norm", "prior_norm"))
            if is_norm:
                p.copy_(torch.ones_like(p) if name.endswith("weight") else torch.zeros_like(p))
                continue
            fan_in = p.shape[1] if (p.dim() == 2 and name.endswith("weight")) else 256
            u = torch.rand(p.shape, generator=g, device="cuda", dtype=torch.float32) * 2 - 1
            p.copy_((u / math.sqrt(fan_in)).to(dtype))
    return model


def algorithmic_flops(cfg, B):
    d = prior_dims(cfg)
    C, S, Lyr = d["inner"], d["seq"], d["layers"]
    rows = B * S
    per_row_layer = 12 * C * C + (22 * C * C if d["motion"] else 0)
    lin = 2.0 * rows * Lyr * per_row_layer
    attn = 4.0 * B * S * S * C * Lyr                     # masked self-attention: QK^T + PV
    tattn = 4.0 * rows * 5 * C * Lyr * d["n_tattn"] if d["motion"] else 0.0
    edge = 2.0 * (B // 2) * d["emb"] * C + 2.0 * B * C * d["clip_dim"]  # proj_in (per step) + proj_to_clip
    return lin + attn + tattn + edge


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--layers", type=int, default=20)
    ap.add_argument("--dtype", default="f16")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    dtype = torch.float16 if a.dtype == "f16" else torch.bfloat16
    cfg = prior_full_config(num_layers=a.layers)
    model = device_random_weights(MyPriorTransformer.from_config(cfg), dtype)
    n_params = sum(p.numel() for p in model.parameters())
    inp = synthetic_prior_inputs(cfg, clip_index=0)
    dev = {k: v.to("cuda", dtype) if v.is_floating_point() else v.cuda() for k, v in inp.items()}
    pipe = Seq_Inpaint_Prior_Pipeline(prior=model, image_encoder=None, text_encoder=None, tokenizer=None,
                                      scheduler=UnCLIPScheduler(**PRIOR_SCHEDULER_KWARGS))
    pipe.use_cuda_graph = not a.no_graph
    g = torch.Generator(device="cuda").manual_seed(42)

    def run(steps):
        return pipe.sample(dev["latents"], dev["prompt_embeds"], dev["text_hidden"], dev["text_mask"],
                           dev["imgs_proj_embeds1"], dev["mask_label"], steps, 4.0, generator=g)

    if a.once:
        pipe.use_cuda_graph = False
        out = run(2)
        torch.cuda.synchronize()
        print("ok", float(out.float().abs().mean()))
        return
    out = run(a.steps)  # warm-up (packs weights, sets kernel attributes)
    torch.cuda.synchronize()
    l0 = _lib.lib().rcdm_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = []
    for _ in range(a.reps):
        e0.record()
        out = run(a.steps)
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    ms = sum(best) / len(best)
    launches = _lib.lib().rcdm_kernel_launches() - l0
    fl = algorithmic_flops(cfg, 10)
    print(json.dumps({
        "metric": f"stage-1 prior frame-embeddings/sec @{a.steps} UnCLIP steps (5-frame clip, CFG 4.0)",
        "value": 5.0 / (ms / 1e3), "unit": "frame-embeddings/s", "ms_per_clip": ms, "ms_per_step": ms / a.steps,
        "steps": a.steps, "reps": a.reps, "dtype": a.dtype, "layers": a.layers, "params_billion": n_params / 1e9,
        "cuda_graph": pipe.use_cuda_graph, "tflop_per_forward": fl / 1e12,
        "achieved_tflops": fl * a.steps / (ms / 1e3) / 1e12,
        "weight_bytes_per_step_gb": n_params * 2 / 1e9,
        "weight_stream_gbs": n_params * 2 / (ms / a.steps / 1e3) / 1e9,
        "launches_recorded_per_clip": launches / a.reps,
        "includes": "whole sampling loop incl. once-per-clip prologue (static tokens, time-embedding table, noise draws)",
        "finite": bool(torch.isfinite(out).all())}))


if __name__ == "__main__":
    main()
