"""Host-side DDIM scheduler with the surface ``RCDMsPipeline`` uses from
``diffusers.DDIMScheduler`` (0.24.0): ``set_timesteps``, ``timesteps``, ``scale_model_input``,
``step(...).prev_sample``, ``init_noise_sigma``, ``order``, ``config``/``_internal_dict``
(reference call sites: ``src/pipelines/RCDMs_pipeline.py:84-109,347,455-456,478,483,497``;
constructed at ``stage2_batchtest_rcdms_model.py:247`` from ``configs/testing.yaml:18-21``).

The schedule tables (betas, alpha-bar, timestep indices) are host integers/fp32 — they are
what must be bit-identical to the reference.  The per-step arithmetic on latents normally
runs in the fused CUDA kernel (``rcdm_ddim_cfg_step``); ``step`` here is the same formula on
torch tensors for API completeness (CPU tensors, callbacks, non-fused use).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


class FrozenConfig(dict):
    """dict with attribute access (the pipeline reads ``scheduler.config.steps_offset``)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


@dataclass
class DDIMStepOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 timestep_spacing: str = "leading"):
        if trained_betas is not None:
            betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if prediction_type != "epsilon":
            raise NotImplementedError("only epsilon prediction is on the RCDMs stage-2 path")
        if timestep_spacing != "leading":
            raise NotImplementedError("only 'leading' timestep spacing is on the RCDMs stage-2 path")
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self._internal_dict = FrozenConfig(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, prediction_type=prediction_type,
            timestep_spacing=timestep_spacing)
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    @property
    def config(self) -> FrozenConfig:
        return self._internal_dict

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than {n_train}")
        self.num_inference_steps = num_inference_steps
        ratio = n_train // num_inference_steps
        ts = np.flip(np.arange(num_inference_steps, dtype=np.int64) * ratio).copy() + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def step_coefficients(self, timestep: int):
        """(alpha_bar_t, alpha_bar_prev) as fp32 python floats for one step — what the fused
        CUDA step consumes.  Index arithmetic identical to diffusers' ``step``."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        prev = int(timestep) - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[int(timestep)]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return float(a_t), float(a_prev)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None,
             return_dict: bool = True):
        if eta != 0.0:
            raise NotImplementedError("eta > 0 (stochastic DDIM) is not on the RCDMs path (eta=0.0 default)")
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        # 0-dim fp32 CPU tensors exactly like diffusers 0.24 (scheduling_ddim.step): the arithmetic on 16-bit latents is
        # then op-for-op what the reference executes (the fused CUDA step mirrors the same rounding points).
        prev = int(timestep) - self.config.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[int(timestep)]
        alpha_prod_t_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        x0 = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-1.0, 1.0)
        pred_sample_direction = (1 - alpha_prod_t_prev) ** 0.5 * model_output
        prev_sample = alpha_prod_t_prev ** 0.5 * x0 + pred_sample_direction
        if not return_dict:
            return (prev_sample,)
        return DDIMStepOutput(prev_sample=prev_sample, pred_original_sample=x0)


# -------------------------------------------------------------------------------------------------------------------
# UnCLIP scheduler (stage-1 frame prior)
# -------------------------------------------------------------------------------------------------------------------
@dataclass
class UnCLIPStepOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class UnCLIPScheduler:
    """Host-side scheduler with the surface ``Seq_Inpaint_Prior_Pipeline`` uses from ``diffusers.UnCLIPScheduler``
    (0.24.0): ``set_timesteps``, ``timesteps``, ``init_noise_sigma``, ``step(pred, timestep=, sample=, generator=,
    prev_timestep=).prev_sample`` (reference call sites: ``src/pipelines/prior_pipeline.py:111,286-287,322-333``;
    constructed at ``stage1_batchtest_rcdms_model.py:101``).

    The tables (cosine betas, alpha-bar, timestep indices) are host fp32/int64.  The per-step arithmetic on the
    embeddings normally runs in the fused CUDA kernel (``rcdm_unclip_cfg_step``) fed by ``step_coefficients``;
    ``step`` is the same formula on torch tensors for API completeness."""

    def __init__(self, num_train_timesteps: int = 1000, variance_type: str = "fixed_small_log",
                 clip_sample: bool = True, clip_sample_range: float = 1.0, prediction_type: str = "epsilon",
                 beta_schedule: str = "squaredcos_cap_v2"):
        if beta_schedule != "squaredcos_cap_v2":
            raise ValueError("UnCLIPScheduler only supports `beta_schedule`: 'squaredcos_cap_v2'")
        if variance_type != "fixed_small_log":
            raise NotImplementedError("variance_type 'learned_range' needs a variance head the RCDMs prior lacks")
        if prediction_type not in ("epsilon", "sample"):
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon` or `sample`")
        # cosine schedule in float64, stored as float32 (diffusers' betas_for_alpha_bar)
        i = np.arange(num_train_timesteps + 1, dtype=np.float64) / num_train_timesteps
        abar = np.cos((i + 0.008) / 1.008 * np.pi / 2) ** 2
        betas = np.minimum(1.0 - abar[1:] / abar[:-1], 0.999)
        self.betas = torch.tensor(betas.tolist(), dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.variance_type = variance_type
        self._internal_dict = FrozenConfig(
            num_train_timesteps=num_train_timesteps, variance_type=variance_type, clip_sample=clip_sample,
            clip_sample_range=clip_sample_range, prediction_type=prediction_type, beta_schedule=beta_schedule)
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    @property
    def config(self) -> FrozenConfig:
        return self._internal_dict

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.num_inference_steps = num_inference_steps
        ratio = (self.config.num_train_timesteps - 1) / (num_inference_steps - 1)
        ts = np.flip((np.arange(num_inference_steps) * ratio).round()).copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)

    def _terms(self, t: int, prev_t: Optional[int]):
        prev_t = t - 1 if prev_t is None else int(prev_t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        if prev_t == t - 1:
            beta, alpha = self.betas[t], self.alphas[t]
        else:
            beta = 1 - a_t / a_prev
            alpha = 1 - beta
        return a_t, a_prev, b_t, b_prev, beta, alpha

    def step_coefficients(self, timestep: int, prev_timestep: Optional[int] = None):
        """(x0_coeff, sample_coeff, sigma, eps_scale, eps_div) as fp32 python floats for one step — what the fused
        CUDA step consumes: x_prev = x0_coeff * clip(x0) + sample_coeff * x + sigma * noise (sigma = 0 at t = 0);
        for epsilon prediction x0 = (x - eps_scale * pred) / eps_div."""
        t = int(timestep)
        a_t, a_prev, b_t, b_prev, beta, alpha = self._terms(t, prev_timestep)
        c_x0 = (a_prev ** 0.5 * beta) / b_t
        c_x = alpha ** 0.5 * b_prev / b_t
        sigma = 0.0
        if t > 0:
            var = b_prev / b_t * beta
            sigma = float(torch.exp(0.5 * torch.log(torch.clamp(var, min=1e-20))))
        return float(c_x0), float(c_x), sigma, float(b_t ** 0.5), float(a_t ** 0.5)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, prev_timestep=None, generator=None,
             return_dict: bool = True):
        t = int(timestep)
        a_t, a_prev, b_t, b_prev, beta, alpha = self._terms(t, prev_timestep)
        if self.config.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        else:
            x0 = model_output
        if self.config.clip_sample:
            x0 = torch.clamp(x0, -self.config.clip_sample_range, self.config.clip_sample_range)
        prev = (a_prev ** 0.5 * beta) / b_t * x0 + alpha ** 0.5 * b_prev / b_t * sample
        if t > 0:
            noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                dtype=model_output.dtype)
            var = b_prev / b_t * beta
            prev = prev + torch.exp(0.5 * torch.log(torch.clamp(var, min=1e-20))) * noise
        if not return_dict:
            return (prev,)
        return UnCLIPStepOutput(prev_sample=prev, pred_original_sample=x0)
