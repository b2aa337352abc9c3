"""Whole-UNet parity helpers (GPU): CUDA path vs the fp32 oracle / the reference golden vectors."""
from __future__ import annotations

import os

import torch

from oracle.unet_ref import unet_forward
from rcdms_b200.models import UNet3DConditionModel
from rcdms_b200.synthetic import synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_inputs(cfg, b, f, h, w, L, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((b, cfg["in_channels"], f, h, w), generator=g)
    ctx = torch.randn((b * f, L, cfg["cross_attention_dim"]), generator=g)
    return x, ctx


def build_model(cfg, dtype, sd=None, simple=False, options=None):
    """Product module with the deterministic synthetic weights, on cuda:0 in `dtype`."""
    sd = sd if sd is not None else synthetic_state_dict(cfg, seed=0)
    m = UNet3DConditionModel.from_config(cfg)
    if simple:
        m.set_debug_option("simple", 1)  # explicit ABI switch (the library reads no environment variables)
    for k, v in (options or {}).items():
        m.set_debug_option(k, v)
    m.load_state_dict(sd, strict=True)
    m = m.to(device="cuda", dtype=dtype)
    return m


def stats(out, ref):
    d = (out.float() - ref.float()).abs()
    return dict(max_abs=d.max().item(), mean_abs=d.mean().item(), ref_mean_abs=ref.float().abs().mean().item(),
                finite=bool(torch.isfinite(out).all()))


def noise_floor(cfg, sd32, x, t, ctx, dtype):
    """|oracle in `dtype` (torch eager on the GPU) - oracle fp32|: the reference's own precision noise."""
    sd16 = {k: v.to(device="cuda", dtype=dtype) for k, v in sd32.items()}
    with torch.no_grad():
        y16 = unet_forward(sd16, cfg, x.to("cuda", dtype), t, ctx.to("cuda", dtype))
    return y16


def run_case(cfg, shape, t, dtype, simple=False, taps=False, sd=None, seed=1234, options=None):
    b, f, h, w, L = shape
    sd = sd if sd is not None else synthetic_state_dict(cfg, seed=0)
    x, ctx = golden_inputs(cfg, b, f, h, w, L, seed)
    m = build_model(cfg, dtype, sd, simple=simple, options=options)
    if taps:
        m.enable_taps(True)
    xd, cd = x.to("cuda", dtype), ctx.to("cuda", dtype)
    y = m(xd, torch.tensor(t, device="cuda"), encoder_hidden_states=cd, return_dict=False)[0]
    torch.cuda.synchronize()
    # fp32 oracle on the GPU fed the same rounded inputs/weights
    sdr = {k: v.to(dtype).to(device="cuda", dtype=torch.float32) for k, v in sd.items()}
    ref_taps = {} if taps else None
    with torch.no_grad():
        ref = unet_forward(sdr, cfg, xd.float(), t, cd.float(), taps=ref_taps)
    res = dict(stats=stats(y, ref), y=y, ref=ref, model=m, tap_stats={})
    if taps:
        for name, rt in ref_taps.items():
            try:
                mine = m.read_tap(name)
            except Exception as e:  # noqa: BLE001
                res["tap_stats"][name] = dict(error=str(e))
                continue
            bb, cc, ff, hh, ww = rt.shape
            rtok = rt.permute(0, 2, 3, 4, 1).reshape(-1, cc)
            res["tap_stats"][name] = stats(mine, rtok)
    del sdr
    y16 = noise_floor(cfg, sd, x, t, ctx, dtype)
    res["floor"] = stats(y16, ref)
    return res
