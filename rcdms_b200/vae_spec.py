"""AutoencoderKL (the SD-1.5 VAE RCDMs loads with ``AutoencoderKL.from_pretrained(..., subfolder="vae")``,
``stage2_batchtest_rcdms_model.py:199``): configuration and the diffusers-0.24 state-dict surface (names + shapes).

diffusers is not vendored in the reference and is absent from this image, so the surface below is RESTATED from the
published diffusers 0.24.0 modules (``models/autoencoder_kl.py``, ``models/vae.py``, ``unet_2d_blocks.py``:
``DownEncoderBlock2D`` / ``UpDecoderBlock2D`` / ``UNetMidBlock2D``, ``resnet.py``: ``ResnetBlock2D`` / ``Downsample2D`` /
``Upsample2D``, ``attention_processor.py``: ``Attention``) - parity unpinned (see oracle/vae_ref.py)."""
from __future__ import annotations

from typing import Dict, List, Tuple

VAE_SD15_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                       layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215)


def vae_full_config(**over) -> Dict:
    c = dict(VAE_SD15_CONFIG)
    c.update(over)
    c["block_out_channels"] = tuple(c["block_out_channels"])
    return c


def vae_tiny_config(**over) -> Dict:
    """Same structure, narrow: 2 levels (4x down / up), one resnet per encoder block."""
    c = dict(VAE_SD15_CONFIG, block_out_channels=(64, 128), layers_per_block=1)
    c.update(over)
    c["block_out_channels"] = tuple(c["block_out_channels"])
    return c


def _resnet(p: str, cin: int, cout: int) -> List[Tuple[str, Tuple[int, ...]]]:
    out = [(p + ".norm1.weight", (cin,)), (p + ".norm1.bias", (cin,)),
           (p + ".conv1.weight", (cout, cin, 3, 3)), (p + ".conv1.bias", (cout,)),
           (p + ".norm2.weight", (cout,)), (p + ".norm2.bias", (cout,)),
           (p + ".conv2.weight", (cout, cout, 3, 3)), (p + ".conv2.bias", (cout,))]
    if cin != cout:
        out += [(p + ".conv_shortcut.weight", (cout, cin, 1, 1)), (p + ".conv_shortcut.bias", (cout,))]
    return out


def _mid(p: str, c: int) -> List[Tuple[str, Tuple[int, ...]]]:
    a = p + ".attentions.0"
    out = [(a + ".group_norm.weight", (c,)), (a + ".group_norm.bias", (c,))]
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out += [(f"{a}.{n}.weight", (c, c)), (f"{a}.{n}.bias", (c,))]
    return out + _resnet(p + ".resnets.0", c, c) + _resnet(p + ".resnets.1", c, c)


def vae_state_dict_spec(cfg: Dict) -> List[Tuple[str, Tuple[int, ...]]]:
    boc, L, lat = cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]
    n = len(boc)
    s: List[Tuple[str, Tuple[int, ...]]] = [("encoder.conv_in.weight", (boc[0], cfg["in_channels"], 3, 3)),
                                            ("encoder.conv_in.bias", (boc[0],))]
    prev = boc[0]
    for i, c in enumerate(boc):
        for j in range(L):
            s += _resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i < n - 1:
            s += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (c, c, 3, 3)),
                  (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (c,))]
        prev = c
    s += _mid("encoder.mid_block", boc[-1])
    s += [("encoder.conv_norm_out.weight", (boc[-1],)), ("encoder.conv_norm_out.bias", (boc[-1],)),
          ("encoder.conv_out.weight", (2 * lat, boc[-1], 3, 3)), ("encoder.conv_out.bias", (2 * lat,))]
    s += [("decoder.conv_in.weight", (boc[-1], lat, 3, 3)), ("decoder.conv_in.bias", (boc[-1],))]
    rev = tuple(reversed(boc))
    prev = rev[0]
    for i, c in enumerate(rev):
        for j in range(L + 1):
            s += _resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i < n - 1:
            s += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (c, c, 3, 3)),
                  (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (c,))]
        prev = c
    s += _mid("decoder.mid_block", boc[-1])
    s += [("decoder.conv_norm_out.weight", (boc[0],)), ("decoder.conv_norm_out.bias", (boc[0],)),
          ("decoder.conv_out.weight", (cfg["out_channels"], boc[0], 3, 3)), ("decoder.conv_out.bias", (cfg["out_channels"],))]
    s += [("quant_conv.weight", (2 * lat, 2 * lat, 1, 1)), ("quant_conv.bias", (2 * lat,)),
          ("post_quant_conv.weight", (lat, lat, 1, 1)), ("post_quant_conv.bias", (lat,))]
    return s


def synthetic_vae_state_dict(cfg: Dict, seed: int = 0):
    """Deterministic random weights with torch's default-init scale (fan-in uniform); norm weights near 1."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in vae_state_dict_spec(cfg):
        if "norm" in name.split(".")[-2]:
            sd[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if name.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
            continue
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        if name.endswith("bias"):
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        else:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * (3.0 / fan_in) ** 0.5
    return sd
