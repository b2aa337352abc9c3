"""Deterministic synthetic weights and PororoSV-shaped inputs.

No pretrained weights, datasets or checkpoints exist offline, so tests and ``bench.py`` run
on synthetic tensors that depend only on (entry name, shape, seed): the reference modules
in the build container (golden generation) and this package on the GPU box therefore see
bit-identical fp32 values without sharing a 5 GB file.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import torch

from .unet_spec import BUFFER_SUFFIX, state_dict_spec


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def positional_encoding(max_len: int, d_model: int) -> torch.Tensor:
    """``PositionalEncoding.pe`` buffer, ``src/models/motion_module.py:256-262``."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def synthetic_tensor(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """fp32 CPU tensor for one state-dict entry.

    Linear/conv weights and all biases: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's default
    scale, so activations stay O(1) through the 100+ layers); norm scales 1 + 0.2 U(-1,1) and
    norm shifts 0.2 U(-1,1) so affine parameters are exercised; ``pe`` from its closed form.
    The zero-initialised temporal ``proj_out`` (``motion_module.py:84-85``) is deliberately
    NOT zero here, otherwise the 20 motion modules would be identities and untested.
    """
    if name.endswith(BUFFER_SUFFIX):
        return positional_encoding(shape[1], shape[2])
    g = _gen(name, seed)
    u = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1
    is_norm = any(k in name for k in (".norm", "norms.", "ff_norm", "prior_norm", "conv_norm_out"))
    if is_norm:
        return 1 + 0.2 * u if name.endswith("weight") else 0.2 * u
    if name.endswith("weight"):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
    else:  # bias: fan-in of the matching weight is unknown here; a fixed small scale is enough
        fan_in = 256
    return u / math.sqrt(fan_in)


def synthetic_state_dict(cfg: Dict, seed: int = 0, dtype=torch.float32, device="cpu",
                         names: Iterable[str] = None) -> Dict[str, torch.Tensor]:
    out = {}
    want = set(names) if names is not None else None
    for name, shape in state_dict_spec(cfg):
        if want is not None and name not in want:
            continue
        out[name] = synthetic_tensor(name, shape, seed).to(device=device, dtype=dtype)
    return out


def synthetic_clip_inputs(clip_index: int, h: int, w: int, ctx_len: int = 85, ctx_dim: int = 768,
                          frames: int = 5, seed: int = 42) -> Dict[str, torch.Tensor]:
    """Per-clip inputs of the denoise loop (fp32, CPU), generator seed ``seed + clip_index``
    (42 = the reference's ``--seed_number`` default, ``stage2_batchtest_rcdms_model.py:426``).

    latents ~ N(0,1) (1,4,f,h,w); masked_latents = 0.18215 N(0,1) (stand-in for the VAE posterior
    sample, ``RCDMs_pipeline.py:429-431``); mask = [1,0,0,0,0] per frame ('continue' mode,
    ``stage2_batchtest_rcdms_model.py:286-288``); ctx ~ N(0,1) (2f, L, ctx_dim).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + clip_index)
    latents = torch.randn((1, 4, frames, h, w), generator=g)
    masked_latents = 0.18215 * torch.randn((1, 4, frames, h, w), generator=g)
    mask = torch.zeros((1, 1, frames, h, w))
    mask[:, :, 0] = 1.0
    ctx = torch.randn((2 * frames, ctx_len, ctx_dim), generator=g)
    return dict(latents=latents, masked_latents=masked_latents, mask=mask, ctx=ctx)


# ---------------------------------------------------------------------------------------------------------------
# stage-1 frame prior
# ---------------------------------------------------------------------------------------------------------------
def synthetic_prior_state_dict(cfg: Dict, seed: int = 0, dtype=torch.float32, device="cpu") -> Dict[str, torch.Tensor]:
    """Deterministic weights for ``MyPriorTransformer`` (same rules as ``synthetic_tensor``; the zero-initialised
    temporal ``proj_out`` is again NOT zero so that the motion modules are exercised).  ``positional_embedding`` /
    ``prd_embedding`` (zeros in the reference's init) get small random values."""
    from .prior_spec import prior_state_dict_spec
    out = {}
    for name, shape in prior_state_dict_spec(cfg):
        out[name] = synthetic_tensor("prior." + name, shape, seed).to(device=device, dtype=dtype)
    return out


def synthetic_prior_inputs(cfg: Dict, clip_index: int = 0, frames: int = 5, seed: int = 42,
                           steps: int = 0) -> Dict[str, torch.Tensor]:
    """Per-clip inputs of the prior sampling loop (fp32, CPU), shaped like ``prior_pipeline.py:283-298`` after
    ``_encode_prompt``: CFG rows ordered [negative x frames, positive x frames].

    latents ~ N(0,1) (f, D); prompt_embeds (2f, D); text_hidden (2f, L, D); text_mask (2f, L) bool — the empty negative
    prompt keeps 2 tokens (BOS/EOS), positive prompts 12..L tokens; imgs_proj_embeds1 / mask_label (f, 1, D) (CLIP image
    embeds of the source frames / of the white-black mask label images, ``stage1_batchtest_rcdms_model.py:166-178``);
    noise (steps-1, f, D): the variance noise of every step but the last (when ``steps`` > 0)."""
    D, L = cfg["embedding_dim"], cfg["num_embeddings"]
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + clip_index)
    latents = torch.randn((frames, D), generator=g)
    prompt_embeds = torch.randn((2 * frames, D), generator=g)
    text_hidden = torch.randn((2 * frames, L, D), generator=g)
    text_mask = torch.zeros((2 * frames, L), dtype=torch.bool)
    text_mask[:frames, :2] = True
    for i in range(frames):
        n = min(L, 12 + (7 * i + 3 * clip_index) % max(1, L - 11))
        text_mask[frames + i, :n] = True
    out = dict(latents=latents, prompt_embeds=prompt_embeds, text_hidden=text_hidden, text_mask=text_mask,
               imgs_proj_embeds1=torch.randn((frames, 1, D), generator=g),
               mask_label=torch.randn((frames, 1, D), generator=g))
    if steps > 0:
        out["noise"] = torch.randn((max(steps - 1, 1), frames, D), generator=g)
    return out


def stack_prior_clips(clips) -> Dict[str, torch.Tensor]:
    """Batch several clips' prior inputs (each as returned by ``synthetic_prior_inputs``) for ONE sampling run: the
    classifier-free-guidance halves stay outermost — rows [negative clip 0..k | positive clip 0..k] — and every clip
    keeps its 5 consecutive frame rows (the prior-state motion modules attend within groups of 5 rows)."""
    f = clips[0]["latents"].shape[0]
    out = dict(latents=torch.cat([c["latents"] for c in clips]),
               imgs_proj_embeds1=torch.cat([c["imgs_proj_embeds1"] for c in clips]),
               mask_label=torch.cat([c["mask_label"] for c in clips]))
    for k in ("prompt_embeds", "text_hidden", "text_mask"):
        out[k] = torch.cat([c[k][:f] for c in clips] + [c[k][f:] for c in clips])
    if "noise" in clips[0]:
        out["noise"] = torch.cat([c["noise"] for c in clips], dim=1)
    return out
