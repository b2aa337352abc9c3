"""Stage-1 pipeline with the reference's operator API (``src/pipelines/prior_pipeline.py:83-373``) whose sampling loop
(``:283-352``) runs on librcdm_b200 kernels: per step [proj_in -> token assembly -> 20 x (masked-attention transformer
block + prior-state motion module) -> norm_out / proj_to_clip -> classifier-free guidance + UnCLIP scheduler step],
captured once into a CUDA graph and replayed for every step (the step index lives in a device counter).

What stays PyTorch (once per clip, outside the loop): prompt encoding with the caller's CLIP text encoder / tokenizer,
``get_zero_embed`` through the caller's CLIP image encoder, and drawing the random numbers.  The noise of every step
is drawn up front from the caller's generator in the order the reference draws it (initial latents first, then one
``randn`` per step with t > 0), so a given generator state yields the same random numbers as the reference loop.

Constructor keywords, ``__call__`` parameters, the error behaviour and the returned ``(image_embeds,
negative_image_embeds)`` mirror the reference.  SURVEY.md §8(f) rank 1.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..models.myprior_transformer import MyPriorTransformer
from ..schedulers import UnCLIPScheduler


@dataclass
class KandinskyPriorPipelineOutput:
    image_embeds: Union[torch.Tensor, np.ndarray]
    negative_image_embeds: Union[torch.Tensor, np.ndarray]


class Seq_Inpaint_Prior_Pipeline:
    model_cpu_offload_seq = "text_encoder->image_encoder->prior"
    _exclude_from_cpu_offload = ["prior"]
    _callback_tensor_inputs = ["latents", "prompt_embeds", "text_encoder_hidden_states", "text_mask"]

    def __init__(self, prior: MyPriorTransformer, image_encoder, text_encoder, tokenizer, scheduler):
        self.prior, self.image_encoder, self.text_encoder = prior, image_encoder, text_encoder
        self.tokenizer, self.scheduler = tokenizer, scheduler
        self._guidance_scale = 4.0
        self._num_timesteps = 0
        self.use_native_loop = True   # False -> python loop over prior()/scheduler.step() (same maths, for debugging)
        self.use_cuda_graph = True
        self.last_gpu_launches = 0  # librcdm kernels executed by the last native sampling run (graph replays included)
        self._native_state = None   # persistent buffers + captured step graph of the last problem signature

    # ---- DiffusionPipeline-like plumbing ----------------------------------------------------------------------
    def _modules(self):
        return [m for m in (self.prior, self.image_encoder, self.text_encoder) if isinstance(m, nn.Module)]

    def to(self, device=None, dtype=None):
        for m in self._modules():
            m.to(device=device, dtype=dtype)
        return self

    @property
    def device(self) -> torch.device:
        for m in self._modules():
            for p in m.parameters():
                return p.device
        return torch.device("cpu")

    @property
    def _execution_device(self) -> torch.device:
        return self.device

    def progress_bar(self, iterable=None, total=None):
        from tqdm.auto import tqdm
        return tqdm(iterable, total=total, disable=getattr(self, "_progress_disabled", True))

    def set_progress_bar_config(self, disable: bool = True, **_):
        self._progress_disabled = disable

    def maybe_free_model_hooks(self):
        return None

    @property
    def do_classifier_free_guidance(self):
        return self._guidance_scale > 1

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    # ---- host-side helpers (once per clip) ------------------------------------------------------------------------
    def prepare_latents(self, shape, dtype, device, generator, latents, scheduler):
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=device, dtype=dtype)
        else:
            if latents.shape != shape:
                raise ValueError(f"Unexpected latents shape, got {latents.shape}, expected {shape}")
            latents = latents.to(device)
        return latents * scheduler.init_noise_sigma

    def get_zero_embed(self, batch_size=1, device=None):
        device = device or self.device
        size = self.image_encoder.config.image_size
        zero_img = torch.zeros(1, 3, size, size).to(device=device, dtype=self.image_encoder.dtype)
        zero_image_emb = self.image_encoder(zero_img)["image_embeds"]
        return zero_image_emb.repeat(batch_size, 1)

    def _text(self, texts, device):
        inp = self.tokenizer(texts, padding="max_length", max_length=self.text_encoder.max_position_embeddings,
                             truncation=False, return_tensors="pt")
        out = self.text_encoder(inp.input_ids.to(device))
        return out.text_embeds, out.last_hidden_state, inp.attention_mask.bool().to(device)

    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt=None):
        batch_size = len(prompt) if isinstance(prompt, list) else 1
        prompt_embeds, hidden, text_mask = self._text(prompt, device)
        prompt_embeds = prompt_embeds.repeat_interleave(num_images_per_prompt, dim=0)
        hidden = hidden.repeat_interleave(num_images_per_prompt, dim=0)
        text_mask = text_mask.repeat_interleave(num_images_per_prompt, dim=0)
        if do_classifier_free_guidance:
            if negative_prompt is None:
                uncond_tokens = [""] * batch_size
            elif type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got "
                                f"{type(negative_prompt)} != {type(prompt)}.")
            elif isinstance(negative_prompt, str):
                uncond_tokens = [negative_prompt]
            elif batch_size != len(negative_prompt):
                raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(negative_prompt)}, but "
                                 f"`prompt`: {prompt} has batch size {batch_size}. Please make sure that passed "
                                 "`negative_prompt` matches the batch size of `prompt`.")
            else:
                uncond_tokens = negative_prompt
            n_embeds, n_hidden, n_mask = self._text(uncond_tokens, device)
            n_embeds = n_embeds.repeat(1, num_images_per_prompt).view(batch_size * num_images_per_prompt, -1)
            seq_len = n_hidden.shape[1]
            n_hidden = n_hidden.repeat(1, num_images_per_prompt, 1).view(batch_size * num_images_per_prompt, seq_len, -1)
            n_mask = n_mask.repeat_interleave(num_images_per_prompt, dim=0)
            prompt_embeds = torch.cat([n_embeds, prompt_embeds])
            hidden = torch.cat([n_hidden, hidden])
            text_mask = torch.cat([n_mask, text_mask])
        return prompt_embeds, hidden, text_mask

    # ---- the sampling loop ---------------------------------------------------------------------------------------------
    def _native_ok(self, latents, callback) -> bool:
        return (self.use_native_loop and isinstance(self.prior, MyPriorTransformer)
                and isinstance(self.scheduler, UnCLIPScheduler) and latents.is_cuda and callback is None
                and latents.dtype in (torch.float16, torch.bfloat16) and latents.dtype == self.prior.dtype)

    def sample(self, latents, prompt_embeds, text_encoder_hidden_states, text_mask, imgs_proj_embeds1, mask_label,
               num_inference_steps: int, guidance_scale: float, generator=None, noise: Optional[torch.Tensor] = None,
               callback_on_step_end: Optional[Callable] = None,
               callback_on_step_end_tensor_inputs: Optional[List[str]] = None) -> torch.Tensor:
        """Loop of ``prior_pipeline.py:299-343`` from the encoded prompt to the final (un-post-processed) latents.
        latents (f, D), already scaled by init_noise_sigma; prompt_embeds (2f|f, D); text_encoder_hidden_states
        (2f|f, L, D); text_mask (2f|f, L); imgs_proj_embeds1 / mask_label (f, 1, D) (duplicated here for CFG like
        ``:297-298``).  ``noise`` (steps-1, f, D), optional: the variance noise of the steps with t > 0 (tests)."""
        self._guidance_scale = guidance_scale
        do_cfg = self.do_classifier_free_guidance
        self.scheduler.set_timesteps(num_inference_steps, device=latents.device)
        timesteps = self.scheduler.timesteps
        self._num_timesteps = len(timesteps)
        p1 = torch.cat([imgs_proj_embeds1] * 2) if do_cfg else imgs_proj_embeds1
        ml = torch.cat([mask_label] * 2) if do_cfg else mask_label
        if self._native_ok(latents, callback_on_step_end):
            return self._sample_native(latents, prompt_embeds, text_encoder_hidden_states, text_mask, p1, ml,
                                       timesteps, guidance_scale, generator, noise)
        for i, t in enumerate(self.progress_bar(timesteps)):
            x = torch.cat([latents] * 2) if do_cfg else latents
            pred = self.prior(x, timestep=t, proj_embedding=prompt_embeds,
                              encoder_hidden_states=text_encoder_hidden_states, proj_embedding1=p1, mask_label=ml,
                              attention_mask=text_mask).predicted_image_embedding
            if do_cfg:
                pu, pt = pred.chunk(2)
                pred = pu + guidance_scale * (pt - pu)
            prev_timestep = None if i + 1 == timesteps.shape[0] else timesteps[i + 1]
            if noise is not None and int(t) > 0:
                g = _FixedNoise(noise[i])
                latents = _step_with_noise(self.scheduler, pred, t, latents, prev_timestep, g)
            else:
                latents = self.scheduler.step(pred, timestep=t, sample=latents, generator=generator,
                                              prev_timestep=prev_timestep).prev_sample
            if callback_on_step_end is not None:
                kw = {k: locals()[k] for k in (callback_on_step_end_tensor_inputs or ["latents"]) if k in locals()}
                outs = callback_on_step_end(self, i, t, kw)
                latents = outs.pop("latents", latents)
                prompt_embeds = outs.pop("prompt_embeds", prompt_embeds)
                text_encoder_hidden_states = outs.pop("text_encoder_hidden_states", text_encoder_hidden_states)
                text_mask = outs.pop("text_mask", text_mask)
        return latents

    def _sample_native(self, latents, prompt_embeds, hidden, text_mask, p1, ml, timesteps, guidance_scale, generator,
                       noise):
        prior, sched = self.prior, self.scheduler
        prior._ensure_packed()
        count = _lib.lib().rcdm_kernel_launches
        c_begin = count()
        dev, dt = latents.device, latents.dtype
        F, D = latents.shape
        do_cfg = guidance_scale > 1
        B = 2 * F if do_cfg else F
        if prompt_embeds.shape[0] != B:
            raise ValueError(f"prompt_embeds has {prompt_embeds.shape[0]} rows, expected {B}")
        ts = [int(t) for t in timesteps.tolist()]
        n = len(ts)
        # scheduler tables: coefficients per step (host fp32 index arithmetic identical to diffusers)
        coef = torch.zeros((n, 8), dtype=torch.float32)
        for i, t in enumerate(ts):
            c = sched.step_coefficients(t, ts[i + 1] if i + 1 < n else None)
            coef[i, :5] = torch.tensor(c, dtype=torch.float32)
            coef[i, 5] = sched.config.clip_sample_range if sched.config.clip_sample else 0.0
            coef[i, 6] = 1.0 if sched.config.prediction_type == "epsilon" else 0.0
        coef = coef.to(dev)
        # variance noise, drawn in the reference's order: one randn(model_output.shape) per step with t > 0
        n_noise = sum(1 for t in ts if t > 0)
        if noise is None:
            noise = torch.stack([torch.randn((F, D), generator=generator, device=dev, dtype=dt)
                                 for _ in range(n_noise)]) if n_noise else torch.zeros((1, F, D), device=dev, dtype=dt)
        noise_tab = torch.zeros((n, F, D), device=dev, dtype=dt)
        j = 0
        for i, t in enumerate(ts):
            if t > 0:
                noise_tab[i] = noise[j].to(device=dev, dtype=dt)
                j += 1
        # step-invariant inputs.  Everything the captured step reads lives in buffers that persist across calls with the
        # same problem signature, so the CUDA graph of one step is captured once per signature and only replayed later.
        key = (B, F, D, n, bool(do_cfg), float(guidance_scale), dt, str(dev), text_mask is not None,
               tuple(prior._versions()))
        st = self._native_state if self._native_state is not None and self._native_state["key"] == key else None
        base = prior.static_tokens(prompt_embeds, hidden, p1, ml)
        temb = prior.time_embedding_table(ts)
        kb = prior.key_bias(text_mask, B)
        if st is None:
            st = dict(key=key, base=base, temb=temb, kb=kb, lat=latents.clone().contiguous(), coef=coef,
                      noise=noise_tab, step=torch.zeros((1,), dtype=torch.int32, device=dev), graph=None)
            self._native_state = st
        else:
            st["base"].copy_(base)
            st["temb"].copy_(temb)
            if kb is not None:
                st["kb"].copy_(kb)
            st["lat"].copy_(latents)
            st["coef"].copy_(coef)
            st["noise"].copy_(noise_tab)
            st["step"].zero_()
        plan = prior._plan(B)
        lat, step = st["lat"], st["step"]
        L = _lib.lib()
        dtid = _lib.torch_dtype_id(dt)

        def one_step():
            pred = prior.run_tokens(plan, st["base"], st["temb"], lat, F, st["kb"], step)
            _lib.check(L.rcdm_unclip_cfg_step(dtid, pred.data_ptr(), lat.data_ptr(), st["noise"].data_ptr(),
                                              st["coef"].data_ptr(), F * D, int(do_cfg), float(guidance_scale),
                                              step.data_ptr(), 1, _lib.current_stream_ptr()))

        host_steps = 0  # times one_step() ran on the host in this call (each issues st["step_launches"] launches)
        if self.use_cuda_graph and n > 1:
            replays = n
            if st["graph"] is None:
                s = torch.cuda.Stream(device=dev)
                s.wait_stream(torch.cuda.current_stream(dev))
                c0 = count()
                with torch.cuda.stream(s):
                    one_step()  # warm-up outside capture: executes step 0 (lazy kernel attributes, allocator)
                st["step_launches"] = int(count() - c0)
                torch.cuda.current_stream(dev).wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    one_step()  # capture only: nothing executes
                st["graph"] = g
                replays = n - 1
                host_steps = 2
            for _ in range(replays):
                st["graph"].replay()
        else:
            c0 = count()
            one_step()
            st["step_launches"] = int(count() - c0)
            for _ in range(n - 1):
                one_step()
            host_steps = n
        # kernels of this library the GPU executed for this clip: the prologue issued from the host + n steps
        prologue = int(count() - c_begin) - host_steps * st["step_launches"]
        self.last_gpu_launches = prologue + n * st["step_launches"]
        return lat.clone()

    @torch.no_grad()
    def __call__(self, prompt, imgs_proj_embeds1, mask_label, video_length: Optional[int], height=None, width=None,
                 num_videos_per_prompt: Optional[int] = 1, negative_prompt=None, num_inference_steps: int = 25,
                 generator=None, latents: Optional[torch.Tensor] = None, guidance_scale: float = 4.0,
                 output_type: Optional[str] = "pt", return_dict: bool = True,
                 callback_on_step_end: Optional[Callable[[int, int, Dict], None]] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"]):
        if negative_prompt is not None:
            prompt = prompt + negative_prompt
            negative_prompt = 2 * negative_prompt
        device = self._execution_device
        batch_size = 1
        self._guidance_scale = guidance_scale
        prompt_embeds, hidden, text_mask = self._encode_prompt(prompt, device, num_videos_per_prompt,
                                                               self.do_classifier_free_guidance, negative_prompt)
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        embedding_dim = self.prior.config.embedding_dim
        latents = self.prepare_latents((batch_size * video_length, embedding_dim), prompt_embeds.dtype, device,
                                       generator, latents, self.scheduler)
        latents = self.sample(latents, prompt_embeds, hidden, text_mask, imgs_proj_embeds1, mask_label,
                              num_inference_steps, guidance_scale, generator=generator,
                              callback_on_step_end=callback_on_step_end,
                              callback_on_step_end_tensor_inputs=callback_on_step_end_tensor_inputs)
        latents = self.prior.post_process_latents(latents)
        image_embeddings = latents
        if negative_prompt is None:
            zero_embeds = self.get_zero_embed(latents.shape[0], device=latents.device)
        else:
            image_embeddings, zero_embeds = image_embeddings.chunk(2)
        self.maybe_free_model_hooks()
        if output_type not in ["pt", "np"]:
            raise ValueError(f"Only the output types `pt` and `np` are supported not output_type={output_type}")
        if output_type == "np":
            image_embeddings = image_embeddings.cpu().numpy()
            zero_embeds = zero_embeds.cpu().numpy()
        if not return_dict:
            return (image_embeddings, zero_embeds)
        return KandinskyPriorPipelineOutput(image_embeds=image_embeddings, negative_image_embeds=zero_embeds)


class _FixedNoise:
    def __init__(self, noise):
        self.noise = noise


def _step_with_noise(sched, pred, t, sample, prev_timestep, fixed: _FixedNoise):
    """Scheduler step with externally supplied variance noise (test hook of the python loop)."""
    c_x0, c_x, sigma, eps_scale, eps_div = sched.step_coefficients(int(t), prev_timestep)
    x0 = pred if sched.config.prediction_type == "sample" else (sample - eps_scale * pred) / eps_div
    if sched.config.clip_sample:
        x0 = torch.clamp(x0, -sched.config.clip_sample_range, sched.config.clip_sample_range)
    return c_x0 * x0 + c_x * sample + sigma * fixed.noise.to(sample)
