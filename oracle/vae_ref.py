"""TEST INFRASTRUCTURE ONLY (oracle) - functional restatement of diffusers 0.24.0 ``AutoencoderKL`` (the SD-1.5 VAE).

PARITY UNPINNED: ``diffusers==0.24.0`` (``requirements.txt:12``) is not vendored under ``/root/reference`` and is absent
from this image, and the reference holds no tests / golden vectors for the VAE.  The algorithm is restated from the
published modules; parity is anchored on the reference's call sites:

* decode  ``RCDMs_pipeline.py:274-287``: ``latents / 0.18215`` -> ``vae.decode(frame).sample`` frame by frame ->
  ``(x / 2 + 0.5).clamp(0, 1)``
* encode  ``RCDMs_pipeline.py:429-431``: ``vae.encode(src).latent_dist.sample(generator) * 0.18215``

Restated pieces (diffusers 0.24.0): ``Encoder`` / ``Decoder`` (models/vae.py): conv_in -> blocks -> UNetMidBlock2D
(resnet, single-head attention over h*w tokens with GroupNorm(32, eps 1e-6) and residual, resnet) -> GroupNorm -> SiLU
-> conv_out; ``ResnetBlock2D`` (GroupNorm eps 1e-6 -> SiLU -> conv3x3 -> GroupNorm -> SiLU -> conv3x3, 1x1
conv_shortcut when the channel count changes, no time embedding); ``Downsample2D(padding=0)``: F.pad (0,1,0,1) then a
stride-2 conv; ``Upsample2D``: nearest 2x then conv3x3; ``DiagonalGaussianDistribution``: mean, logvar = chunk(2),
logvar clamped to [-30, 20], sample = mean + exp(0.5 logvar) * randn(generator)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
EPS = 1e-6


def _gn(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], EPS)


def _conv(sd: SD, p: str, x: torch.Tensor, stride: int = 1, padding: int = 1) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _resnet(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, groups)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, groups)))
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h  # output_scale_factor = 1


def _attention(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """``Attention(heads=1, residual_connection=True, norm_num_groups=32)`` as the VAE mid block builds it."""
    b, c, h, w = x.shape
    y = _gn(sd, p + ".group_norm", x.reshape(b, c, h * w), groups).transpose(1, 2)  # (b, hw, c)
    q = F.linear(y, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(y, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(y, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    a = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * (c ** -0.5), dim=-1)
    o = F.linear(torch.bmm(a, v), sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, h, w)


def _mid(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    x = _resnet(sd, p + ".resnets.0", x, groups)
    x = _attention(sd, p + ".attentions.0", x, groups)
    return _resnet(sd, p + ".resnets.1", x, groups)


def vae_decode(sd: SD, cfg: Dict, z: torch.Tensor) -> torch.Tensor:
    """``AutoencoderKL.decode(z).sample``: z (n, latent, h, w) -> image (n, 3, 8h, 8w)."""
    g, boc, L = cfg["norm_num_groups"], cfg["block_out_channels"], cfg["layers_per_block"]
    x = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = _conv(sd, "decoder.conv_in", x)
    x = _mid(sd, "decoder.mid_block", x, g)
    for i in range(len(boc)):
        for j in range(L + 1):
            x = _resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x, g)
        if i < len(boc) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(_gn(sd, "decoder.conv_norm_out", x, g))
    return _conv(sd, "decoder.conv_out", x)


def vae_encode_moments(sd: SD, cfg: Dict, x: torch.Tensor) -> torch.Tensor:
    """``quant_conv(encoder(x))``: image (n, 3, H, W) -> moments (n, 2 * latent, H / 8, W / 8)."""
    g, boc, L = cfg["norm_num_groups"], cfg["block_out_channels"], cfg["layers_per_block"]
    x = _conv(sd, "encoder.conv_in", x)
    for i in range(len(boc)):
        for j in range(L):
            x = _resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", x, g)
        if i < len(boc) - 1:
            x = F.pad(x, (0, 1, 0, 1))
            x = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", x, stride=2, padding=0)
    x = _mid(sd, "encoder.mid_block", x, g)
    x = F.silu(_gn(sd, "encoder.conv_norm_out", x, g))
    x = _conv(sd, "encoder.conv_out", x)
    return F.conv2d(x, sd["quant_conv.weight"], sd["quant_conv.bias"])


def gaussian_sample(moments: torch.Tensor, generator: Optional[torch.Generator] = None,
                    noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    mean, logvar = moments.chunk(2, dim=1)
    std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
    if noise is None:
        noise = torch.randn(mean.shape, generator=generator, device=mean.device, dtype=mean.dtype)
    return mean + std * noise
