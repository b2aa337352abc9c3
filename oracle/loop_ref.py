"""TEST INFRASTRUCTURE ONLY (oracle) — restated denoise loop.

Follows ``RCDMsPipeline.__call__`` ``src/pipelines/RCDMs_pipeline.py:476-503`` from the
prepared per-clip tensors (latents, mask, masked latents, context) to the final latents,
with the restated DDIM scheduler configured the way the pipeline forces it
(steps_offset=1, clip_sample=False: ``RCDMs_pipeline.py:84-109``; betas from
``configs/testing.yaml:18-21``).
"""
from __future__ import annotations

from typing import Callable, Dict

import torch

from rcdms_b200.unet_spec import RCDMS_SCHEDULER_KWARGS
from .diffusers_restated import DDIMSchedulerRef
from .unet_ref import unet_forward


def make_scheduler() -> DDIMSchedulerRef:
    return DDIMSchedulerRef(steps_offset=1, clip_sample=False, **RCDMS_SCHEDULER_KWARGS)


def denoise_loop(unet: Callable, latents: torch.Tensor, mask: torch.Tensor, masked_latents: torch.Tensor,
                 ctx: torch.Tensor, num_inference_steps: int, guidance_scale: float = 2.0) -> torch.Tensor:
    """``unet(x, t, ctx) -> eps``; tensors are for ONE clip: latents (1,4,f,h,w), mask (1,1,f,h,w),
    masked_latents (1,4,f,h,w), ctx (2f, L, D).  CFG is on when guidance_scale > 1 (:416)."""
    sched = make_scheduler()
    sched.set_timesteps(num_inference_steps)
    cfg_on = guidance_scale > 1.0
    mask2 = torch.cat([mask] * 2) if cfg_on else mask
    ml2 = torch.cat([masked_latents] * 2) if cfg_on else masked_latents
    for t in sched.timesteps:
        lmi = torch.cat([latents] * 2) if cfg_on else latents            # :482
        lmi = sched.scale_model_input(lmi, t)                            # :483
        x = torch.cat([lmi, mask2, ml2], dim=1).to(latents.dtype)        # :486
        eps = unet(x, t, ctx)                                            # :488
        if cfg_on:
            eu, ec = eps.chunk(2)                                        # :493
            eps = eu + guidance_scale * (ec - eu)                        # :494
        latents = sched.step(eps, t, latents, eta=0.0).prev_sample       # :497
    return latents


def denoise_loop_oracle(sd: Dict[str, torch.Tensor], cfg: Dict, **kw) -> torch.Tensor:
    return denoise_loop(lambda x, t, c: unet_forward(sd, cfg, x, t, c), **kw)
