"""TEST INFRASTRUCTURE ONLY (oracle) — functional fp32 restatement of the reference UNet.

Follows, line by line, the reference forward on a plain ``{name: tensor}`` state dict:

* ``UNet3DConditionModel.forward``        ``src/models/unet.py:322-462``
* block containers                        ``src/models/unet_blocks.py:272-280,384-427,499-531,631-680,748-777``
* ``ResnetBlock3D.forward``               ``src/models/resnet.py:182-212``
* ``InflatedConv3d`` / samplers           ``src/models/resnet.py:10-18,32-106``
* ``Transformer3DModel.forward``          ``src/models/attention.py:318-365``
* ``BasicTransformerBlock.forward``       ``src/models/attention.py:479-526``
* ``CrossAttention.forward/_attention``   ``src/models/attention.py:113-199``
* ``TemporalTransformer3DModel.forward``  ``src/models/motion_module.py:147-182``
* ``TemporalTransformerBlock`` / ``VersatileAttention`` / ``PositionalEncoding``
                                          ``src/models/motion_module.py:234-246,294-354,249-267``
* ``Timesteps`` / ``TimestepEmbedding`` / ``FeedForward`` — restated diffusers 0.24.0
  (``oracle/diffusers_restated.py``).

Pinned against the reference modules themselves (imported unmodified through
``oracle/diffusers_shim``) by the golden vectors in ``tests/golden/`` — see
``oracle/make_golden.py`` and ``tests/test_oracle_golden.py``.

Runs in whatever dtype/device the state dict is in (fp32 CPU for parity; fp32 CUDA is allowed
for full-size checks where the CPU would take minutes).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from rcdms_b200.unet_spec import block_plan
from .diffusers_restated import get_timestep_embedding

SD = Dict[str, torch.Tensor]


def _conv(sd: SD, p: str, x5: torch.Tensor, stride: int = 1, padding: int = 1) -> torch.Tensor:
    """InflatedConv3d: per-frame conv2d on (b c f h w)  — resnet.py:10-18."""
    b, c, f, h, w = x5.shape
    x = x5.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)
    return x.reshape(b, f, *x.shape[1:]).permute(0, 2, 1, 3, 4)


def _resnet(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    """ResnetBlock3D.forward — resnet.py:182-212.  GroupNorm is on the 5-D tensor => statistics
    span all frames (use_inflated_groupnorm=False)."""
    h = F.group_norm(x, groups, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    h = F.silu(h)
    h = _conv(sd, p + ".conv1", h)
    t = F.linear(F.silu(emb), sd[p + ".time_emb_proj.weight"], sd[p + ".time_emb_proj.bias"])
    h = h + t[:, :, None, None, None]
    h = F.group_norm(h, groups, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = F.silu(h)
    h = _conv(sd, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h  # output_scale_factor == 1


def _heads_to_batch(t: torch.Tensor, heads: int) -> torch.Tensor:
    b, s, d = t.shape
    return t.reshape(b, s, heads, d // heads).permute(0, 2, 1, 3).reshape(b * heads, s, d // heads)


def _batch_to_heads(t: torch.Tensor, heads: int) -> torch.Tensor:
    bh, s, d = t.shape
    return t.reshape(bh // heads, heads, s, d).permute(0, 2, 1, 3).reshape(bh // heads, s, d * heads)


def _attention(sd: SD, p: str, x: torch.Tensor, ctx: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """CrossAttention.forward + _attention — attention.py:113-199 (no bias on q/k/v, bias on out)."""
    kv = x if ctx is None else ctx
    q = _heads_to_batch(F.linear(x, sd[p + ".to_q.weight"]), heads)
    k = _heads_to_batch(F.linear(kv, sd[p + ".to_k.weight"]), heads)
    v = _heads_to_batch(F.linear(kv, sd[p + ".to_v.weight"]), heads)
    scale = q.shape[-1] ** -0.5
    scores = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
                           q, k.transpose(-1, -2), beta=0, alpha=scale)
    probs = scores.softmax(dim=-1)
    o = _batch_to_heads(torch.bmm(probs, v), heads)
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _ff(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """diffusers FeedForward with GEGLU — attention.py:434."""
    hg = F.linear(x, sd[p + ".net.0.proj.weight"], sd[p + ".net.0.proj.bias"])
    h, g = hg.chunk(2, dim=-1)
    return F.linear(h * F.gelu(g), sd[p + ".net.2.weight"], sd[p + ".net.2.bias"])


def _ln(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _spatial_transformer(sd: SD, p: str, x5: torch.Tensor, ctx: torch.Tensor, heads: int, groups: int) -> torch.Tensor:
    """Transformer3DModel.forward + BasicTransformerBlock.forward — attention.py:318-365,479-526."""
    b, c, f, h, w = x5.shape
    x = x5.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    res = x
    y = F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    y = F.conv2d(y, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    y = y.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    t = p + ".transformer_blocks.0"
    y = _attention(sd, t + ".attn1", _ln(sd, t + ".norm1", y), None, heads) + y
    y = _attention(sd, t + ".attn2", _ln(sd, t + ".norm2", y), ctx, heads) + y
    y = _ff(sd, t + ".ff", _ln(sd, t + ".norm3", y)) + y
    y = y.reshape(b * f, h, w, c).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    y = y + res
    return y.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def _temporal_attention(sd: SD, p: str, x: torch.Tensor, f: int, heads: int) -> torch.Tensor:
    """VersatileAttention.forward — motion_module.py:294-354: (b f) d c -> (b d) f c, add PE, attend over f."""
    bf, d, c = x.shape
    b = bf // f
    y = x.reshape(b, f, d, c).permute(0, 2, 1, 3).reshape(b * d, f, c)
    y = y + sd[p + ".pos_encoder.pe"][:, :f]
    y = _attention(sd, p, y, None, heads)
    return y.reshape(b, d, f, c).permute(0, 2, 1, 3).reshape(bf, d, c)


def _motion_module(sd: SD, p: str, x5: torch.Tensor, heads: int, groups: int, n_attn: int) -> torch.Tensor:
    """TemporalTransformer3DModel.forward + TemporalTransformerBlock.forward —
    motion_module.py:147-182,234-246 (prior_state=False branch)."""
    p = p + ".temporal_transformer"
    b, c, f, h, w = x5.shape
    x = x5.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    res = x
    y = F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    y = F.linear(y, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    t = p + ".transformer_blocks.0"
    for i in range(n_attn):
        y = _temporal_attention(sd, f"{t}.attention_blocks.{i}", _ln(sd, f"{t}.norms.{i}", y), f, heads) + y
    y = _ff(sd, t + ".ff", _ln(sd, t + ".ff_norm", y)) + y
    y = F.linear(y, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    y = y.reshape(b * f, h, w, c).permute(0, 3, 1, 2)
    y = y + res
    return y.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def time_embedding(sd: SD, cfg: Dict, timestep, batch: int, dtype, device) -> torch.Tensor:
    """unet.py:367-389: sinusoid in fp32, cast to model dtype, linear_1 -> SiLU -> linear_2."""
    t = torch.as_tensor(timestep, device=device).reshape(-1).expand(batch)
    c0 = cfg["block_out_channels"][0]
    t_emb = get_timestep_embedding(t, c0, cfg["flip_sin_to_cos"], cfg["freq_shift"]).to(dtype)
    e = F.linear(t_emb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    return F.linear(F.silu(e), sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])


def unet_forward(sd: SD, cfg: Dict, sample: torch.Tensor, timestep, ctx: torch.Tensor,
                 taps: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """UNet3DConditionModel.forward — unet.py:322-462.  ``taps`` (optional dict) receives
    intermediate activations keyed by block name, for layer-by-layer parity bisection."""
    pl = block_plan(cfg)
    heads, groups, eps = pl["heads"], pl["groups"], cfg["norm_eps"]
    mh, n_t = pl["motion_heads"], pl["n_tattn"]
    any_w = sd["conv_in.weight"]
    emb = time_embedding(sd, cfg, timestep, sample.shape[0], any_w.dtype, any_w.device)

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    x = _conv(sd, "conv_in", sample)
    tap("conv_in", x)
    skips = [x]
    for i, b in enumerate(pl["down"]):
        p = f"down_blocks.{i}"
        for j in range(b["layers"]):
            x = _resnet(sd, f"{p}.resnets.{j}", x, emb, groups, eps)
            tap(f"{p}.resnets.{j}", x)
            if b["attn"]:
                x = _spatial_transformer(sd, f"{p}.attentions.{j}", x, ctx, heads, groups)
                tap(f"{p}.attentions.{j}", x)
            if b["motion"]:
                x = _motion_module(sd, f"{p}.motion_modules.{j}", x, mh, groups, n_t)
                tap(f"{p}.motion_modules.{j}", x)
            skips.append(x)
        if b["sampler"]:
            x = _conv(sd, f"{p}.downsamplers.0.conv", x, stride=2, padding=1)
            tap(f"{p}.downsamplers.0", x)
            skips.append(x)
    x = _resnet(sd, "mid_block.resnets.0", x, emb, groups, eps)
    x = _spatial_transformer(sd, "mid_block.attentions.0", x, ctx, heads, groups)
    if pl["mid_motion"]:
        x = _motion_module(sd, "mid_block.motion_modules.0", x, mh, groups, n_t)
    x = _resnet(sd, "mid_block.resnets.1", x, emb, groups, eps)
    tap("mid_block", x)
    for i, b in enumerate(pl["up"]):
        p = f"up_blocks.{i}"
        for j in range(len(b["layers"])):
            x = torch.cat([x, skips.pop()], dim=1)
            x = _resnet(sd, f"{p}.resnets.{j}", x, emb, groups, eps)
            tap(f"{p}.resnets.{j}", x)
            if b["attn"]:
                x = _spatial_transformer(sd, f"{p}.attentions.{j}", x, ctx, heads, groups)
                tap(f"{p}.attentions.{j}", x)
            if b["motion"]:
                x = _motion_module(sd, f"{p}.motion_modules.{j}", x, mh, groups, n_t)
                tap(f"{p}.motion_modules.{j}", x)
        if b["sampler"]:
            x = F.interpolate(x, scale_factor=[1.0, 2.0, 2.0], mode="nearest")  # resnet.py:65
            x = _conv(sd, f"{p}.upsamplers.0.conv", x)
            tap(f"{p}.upsamplers.0", x)
    x = F.group_norm(x, groups, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], eps)
    x = F.silu(x)
    return _conv(sd, "conv_out", x)
