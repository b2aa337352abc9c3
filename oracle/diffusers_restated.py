"""TEST INFRASTRUCTURE ONLY (oracle) — restated diffusers==0.24.0 arithmetic.

The reference (muzishen/RCDMs) pins ``diffusers==0.24.0`` (``requirements.txt:12``) and
uses, on the stage-2 hot path:

* ``DDIMScheduler``      — constructed ``stage2_batchtest_rcdms_model.py:247`` with
  ``configs/testing.yaml:18-21``; used ``src/pipelines/RCDMs_pipeline.py:84-109,347,455-456,
  478,483,497``.
* ``Timesteps`` / ``TimestepEmbedding`` — ``src/models/unet.py:17,100-103,383-389``.
* ``FeedForward`` (+ ``GEGLU``) — ``src/models/attention.py:14,434``;
  ``src/models/motion_module.py:14,231``.

diffusers is absent from ``/root/reference`` and from this image, so the published
algorithm is restated here.  **Parity unpinned** against a real diffusers install; the
closed-form known answers of SURVEY.md §8(c) pin the schedule, the index arithmetic and
one step (``tests/test_scheduler_known_answers.py``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# Timestep sinusoid  (diffusers.models.embeddings.get_timestep_embedding / Timesteps)
# ----------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, flip_sin_to_cos: bool = False,
                           downscale_freq_shift: float = 1, scale: float = 1, max_period: int = 10000):
    """emb_i = t * exp(-ln(max_period) * i / (half - shift)); cat[sin, cos]; swap halves if flip."""
    assert timesteps.dim() == 1
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    """linear_1 -> SiLU -> linear_2 (only the arguments the reference uses)."""

    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu", out_dim: Optional[int] = None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        assert act_fn == "silu"
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


# ----------------------------------------------------------------------------------------
# FeedForward / GEGLU  (diffusers.models.attention)
# ----------------------------------------------------------------------------------------
class GELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return F.gelu(self.proj(x))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)  # erf GELU


class FeedForward(nn.Module):
    """net = [GEGLU|GELU(dim, 4 dim), Dropout, Linear(4 dim, dim)] — parameter names
    ``net.0.proj.{weight,bias}``, ``net.2.{weight,bias}`` as in the reference state dict."""

    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, dropout: float = 0.0,
                 activation_fn: str = "geglu", final_dropout: bool = False):
        super().__init__()
        inner = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        if activation_fn == "gelu":
            act = GELU(dim, inner)
        elif activation_fn == "geglu":
            act = GEGLU(dim, inner)
        else:
            raise ValueError(activation_fn)
        self.net = nn.ModuleList([act, nn.Dropout(dropout), nn.Linear(inner, dim_out)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class AdaLayerNorm(nn.Module):  # imported by the reference, never instantiated on this path
    def __init__(self, *a, **k):
        raise NotImplementedError("AdaLayerNorm is not on the stage-2 path (num_embeds_ada_norm=None)")


# ----------------------------------------------------------------------------------------
# DDIM scheduler (diffusers.schedulers.scheduling_ddim.DDIMScheduler, 0.24.0 semantics)
# ----------------------------------------------------------------------------------------
@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class DDIMSchedulerRef:
    """Restated DDIMScheduler.  Defaults are diffusers' (steps_offset=0, clip_sample=True);
    the reference pipeline overrides them to 1 / False at construction
    (``RCDMs_pipeline.py:84-109``), which callers of this oracle must mirror."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 timestep_spacing="leading"):
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           beta_schedule=beta_schedule, clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one,
                           steps_offset=steps_offset, prediction_type=prediction_type,
                           timestep_spacing=timestep_spacing)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        assert self.config.timestep_spacing == "leading"
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        timesteps += self.config.steps_offset
        self.timesteps = torch.from_numpy(timesteps).to(device)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output=False,
             generator=None, variance_noise=None, return_dict=True):
        assert eta == 0.0, "oracle covers eta=0 (the reference default, RCDMs_pipeline.py:385)"
        timestep = int(timestep)
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        assert self.config.prediction_type == "epsilon"
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        pred_epsilon = model_output
        if self.config.clip_sample:
            pred_original_sample = pred_original_sample.clamp(-1.0, 1.0)
        # eta == 0 -> variance 0, std_dev_t 0
        pred_sample_direction = (1 - alpha_prod_t_prev) ** 0.5 * pred_epsilon
        prev_sample = alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction
        return DDIMSchedulerOutput(prev_sample=prev_sample, pred_original_sample=pred_original_sample)


# ----------------------------------------------------------------------------------------
# UnCLIP scheduler (diffusers.schedulers.scheduling_unclip.UnCLIPScheduler, 0.24.0 semantics)
# used by the stage-1 prior: stage1_batchtest_rcdms_model.py:31,101; src/pipelines/prior_pipeline.py:286-287,
# 330-336.  Restated from the published algorithm — parity unpinned (see module header); pinned by the
# closed-form known answers in tests/test_scheduler_known_answers.py.
# ----------------------------------------------------------------------------------------
def betas_for_alpha_bar(num_diffusion_timesteps: int, max_beta: float = 0.999) -> torch.Tensor:
    """squaredcos_cap_v2: alpha_bar(t) = cos((t + 0.008) / 1.008 * pi / 2)^2, beta_i = min(1 - ab(t2)/ab(t1), max_beta),
    computed in Python doubles and stored as float32."""
    def alpha_bar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return torch.tensor(betas, dtype=torch.float32)


@dataclass
class UnCLIPSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class UnCLIPSchedulerRef:
    """Restated UnCLIPScheduler (DDPM variant with an explicit ``prev_timestep``)."""

    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, variance_type="fixed_small_log", clip_sample=True,
                 clip_sample_range=1.0, prediction_type="epsilon", beta_schedule="squaredcos_cap_v2"):
        if beta_schedule != "squaredcos_cap_v2":
            raise ValueError("UnCLIPScheduler only supports `beta_schedule`: 'squaredcos_cap_v2'")
        self.betas = betas_for_alpha_bar(num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self.variance_type = variance_type
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, variance_type=variance_type,
                           clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                           prediction_type=prediction_type, beta_schedule=beta_schedule)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = (self.config.num_train_timesteps - 1) / (self.num_inference_steps - 1)
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(timesteps).to(device)

    def _get_variance(self, t, prev_timestep=None, predicted_variance=None, variance_type=None):
        if prev_timestep is None:
            prev_timestep = t - 1
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        beta = self.betas[t] if prev_timestep == t - 1 else 1 - alpha_prod_t / alpha_prod_t_prev
        variance = beta_prod_t_prev / beta_prod_t * beta
        variance_type = variance_type or self.config.variance_type
        if variance_type == "fixed_small_log":
            variance = torch.log(torch.clamp(variance, min=1e-20))
            variance = torch.exp(0.5 * variance)
        else:
            raise NotImplementedError("learned_range needs a variance head the RCDMs prior does not have")
        return variance

    def step(self, model_output, timestep, sample, prev_timestep=None, generator=None, return_dict=True,
             variance_noise=None):
        """``variance_noise`` (oracle-only extension): use this tensor instead of drawing from ``generator`` so that
        two implementations can be compared on identical noise."""
        t = int(timestep)
        prev_timestep = t - 1 if prev_timestep is None else int(prev_timestep)
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        if prev_timestep == t - 1:
            beta = self.betas[t]
            alpha = self.alphas[t]
        else:
            beta = 1 - alpha_prod_t / alpha_prod_t_prev
            alpha = 1 - beta
        if self.config.prediction_type == "epsilon":
            pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        elif self.config.prediction_type == "sample":
            pred_original_sample = model_output
        else:
            raise ValueError(self.config.prediction_type)
        if self.config.clip_sample:
            pred_original_sample = torch.clamp(pred_original_sample, -self.config.clip_sample_range,
                                               self.config.clip_sample_range)
        pred_original_sample_coeff = (alpha_prod_t_prev ** 0.5 * beta) / beta_prod_t
        current_sample_coeff = alpha ** 0.5 * beta_prod_t_prev / beta_prod_t
        pred_prev_sample = pred_original_sample_coeff * pred_original_sample + current_sample_coeff * sample
        variance = 0
        if t > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype,
                                             device=model_output.device)
            variance = self._get_variance(t, prev_timestep=prev_timestep) * variance_noise
        pred_prev_sample = pred_prev_sample + variance
        return UnCLIPSchedulerOutput(prev_sample=pred_prev_sample, pred_original_sample=pred_original_sample)
