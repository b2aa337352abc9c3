"""Stage-2 pipeline with the reference's operator API (``src/pipelines/RCDMs_pipeline.py:61-517``) whose
denoise loop (``:476-503``) runs natively: one call to ``rcdm_denoise_loop`` replays a CUDA graph of
[UNet forward -> classifier-free guidance -> DDIM step -> next 9-channel input] for every step.

What stays PyTorch (outside the hot loop): prompt encoding with the caller's CLIP text
encoder/tokenizer, VAE encode of the reference frames / decode of the result, and the two context-fusion
modules (``local_feature``).  Constructor keywords, ``__call__`` parameters, error behaviour (``ValueError`` /
``TypeError`` for the same conditions) and the returned ``RCDMsPipelineOutput.videos`` (1,3,f,H,W) cpu fp32 in
[0,1] mirror the reference.  Two generalisations that are identities at the reference's only operating point
(512x512, one clip): the mask view uses the actual latent size instead of the hard-coded ``(2,1,5,64,64)``
(``:476``), and ``encode_mask``'s frame count follows ``video_length`` instead of the literal 5 (``:261``).
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..models.unet import UNet3DConditionModel
from ..schedulers import DDIMScheduler, FrozenConfig


class local_feature(nn.Module):
    """Context fusion (reference ``RCDMs_pipeline.py:35-55`` = ``fine_stack`` / ``semantic_stack`` of
    ``stage2_batchtest_rcdms_model.py:117-149``): text tokens query the visual tokens through one
    ``nn.MultiheadAttention`` (seq-first), after a Linear on each side.  Parameters are the reference's own torch modules
    (same state-dict names: ``text_fc``, ``vis_fc``, ``multihead_attn.in_proj_weight / in_proj_bias / out_proj``), so the
    DeepSpeed checkpoint split loads unchanged.

    On a CUDA device in float16 / bfloat16 the forward is six C-ABI launches of ``librcdm_b200``: four tcgen05 GEMMs
    (text_fc, vis_fc, the q and the fused k|v in-projections), ``rcdm_flash_attn`` (8 heads, d = 96, 257 or 1 keys) and the
    out-projection GEMM.  Anywhere else it raises unless ``allow_torch_path`` is set (CPU-tier host-logic tests only)."""

    allow_torch_path = False

    def __init__(self, text_dim: int, vis_dim: int, hidden_dim: int = 768, num_heads: int = 8):
        super().__init__()
        self.hidden_dim, self.num_heads = hidden_dim, num_heads
        self.text_fc = nn.Linear(text_dim, hidden_dim)
        self.vis_fc = nn.Linear(vis_dim, hidden_dim)
        self.multihead_attn = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=num_heads)  # (seq, batch, dim)

    def forward_torch(self, vis_f: torch.Tensor, text_f: torch.Tensor) -> torch.Tensor:
        """The reference's op sequence in torch (``stage2_batchtest_rcdms_model.py:142-147``)."""
        q = self.text_fc(text_f).transpose(0, 1)
        kv = self.vis_fc(vis_f).transpose(0, 1)
        return self.multihead_attn(q, kv, kv)[0].transpose(0, 1)

    @staticmethod
    def _lin(a: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        from .. import ops
        if a.shape[1] % 8:  # the TMA rows need 16-byte multiples: zero-pad K (plumbing; e.g. 12-wide test embeddings)
            pad = 8 - a.shape[1] % 8
            a = torch.nn.functional.pad(a, (0, pad))
            w = torch.nn.functional.pad(w, (0, pad))
        return ops.linear(a.contiguous(), w.contiguous(), b.float().contiguous())

    @torch.no_grad()
    def forward(self, vis_f: torch.Tensor, text_f: torch.Tensor) -> torch.Tensor:
        native = text_f.is_cuda and vis_f.is_cuda and text_f.dtype in (torch.float16, torch.bfloat16) \
            and self.text_fc.weight.dtype == text_f.dtype
        if not native:
            if not self.allow_torch_path:
                raise RuntimeError("local_feature runs on the B200 kernels: CUDA tensors and a float16 / bfloat16 module "
                                   "(set allow_torch_path = True for the reference's torch ops in host-logic tests)")
            return self.forward_torch(vis_f, text_f)
        from .. import _lib
        B, L, _ = text_f.shape
        Sv, D, H = vis_f.shape[1], self.hidden_dim, self.num_heads
        vis_f = vis_f.to(text_f.dtype)
        w_in, b_in = self.multihead_attn.in_proj_weight, self.multihead_attn.in_proj_bias
        q = self._lin(self._lin(text_f.reshape(B * L, -1), self.text_fc.weight, self.text_fc.bias), w_in[:D], b_in[:D])
        kv = self._lin(self._lin(vis_f.reshape(B * Sv, -1), self.vis_fc.weight, self.vis_fc.bias), w_in[D:], b_in[D:])
        d = D // H
        att = torch.empty_like(q)
        simple = 0 if (d % 8 == 0 and d <= 160) else 1  # odd head dims (tiny test modules): the library's CUDA-core kernel
        _lib.check(_lib.lib().rcdm_flash_attn(_lib.torch_dtype_id(q.dtype), q.data_ptr(), D, kv.data_ptr(),
                                              kv.data_ptr() + D * kv.element_size(), 2 * D, att.data_ptr(), D, B, H, L, Sv, d,
                                              simple, _lib.current_stream_ptr()))
        out = self._lin(att, self.multihead_attn.out_proj.weight, self.multihead_attn.out_proj.bias)
        return out.reshape(B, L, D)


@dataclass
class RCDMsPipelineOutput:
    videos: Union[torch.Tensor, np.ndarray]


class RCDMsPipeline:
    _optional_components: List[str] = []

    def __init__(self, vae, text_encoder, tokenizer, unet: UNet3DConditionModel, local_module, global_module,
                 scheduler):
        # the reference patches outdated scheduler configs in place (RCDMs_pipeline.py:84-109)
        cfg = getattr(scheduler, "config", None)
        if cfg is not None and getattr(cfg, "steps_offset", 1) != 1:
            scheduler._internal_dict = FrozenConfig({**dict(cfg), "steps_offset": 1})
        cfg = getattr(scheduler, "config", None)
        if cfg is not None and getattr(cfg, "clip_sample", False) is True:
            scheduler._internal_dict = FrozenConfig({**dict(cfg), "clip_sample": False})
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.unet, self.local_module, self.global_module, self.scheduler = unet, local_module, global_module, scheduler
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)
        self.use_native_loop = True   # False -> python loop over unet()/scheduler.step() (same maths, for debugging)
        self.use_cuda_graph = True

    # ---- DiffusionPipeline-like plumbing -------------------------------------------------------------------
    def _modules(self):
        return [m for m in (self.vae, self.text_encoder, self.unet, self.local_module, self.global_module)
                if isinstance(m, nn.Module)]

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

    def enable_vae_slicing(self):
        self.vae.enable_slicing()

    def disable_vae_slicing(self):
        self.vae.disable_slicing()

    # ---- prompt / mask / latent helpers (host side, once per clip) -------------------------------------------
    def _tokenize_encode(self, texts, device):
        ids = self.tokenizer(texts, padding="max_length", max_length=self.text_encoder.max_position_embeddings,
                             truncation=False, return_tensors="pt").input_ids
        return self.text_encoder(ids.to(device)).last_hidden_state

    def _encode_prompt(self, prompt, device, num_videos_per_prompt, do_classifier_free_guidance, negative_prompt):
        batch_size = len(prompt) if isinstance(prompt, list) else 1
        text_embeddings = self._tokenize_encode(prompt, device)
        if not do_classifier_free_guidance:
            return text_embeddings
        if negative_prompt is None:
            uncond_tokens = [""] * batch_size
        elif type(prompt) is not type(negative_prompt):
            raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} !="
                            f" {type(prompt)}.")
        elif isinstance(negative_prompt, str):
            uncond_tokens = [negative_prompt]
        elif batch_size != len(negative_prompt):
            raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(negative_prompt)}, but `prompt`:"
                             f" {prompt} has batch size {batch_size}. Please make sure that passed `negative_prompt`"
                             " matches the batch size of `prompt`.")
        else:
            uncond_tokens = negative_prompt
        uncond = self._tokenize_encode(uncond_tokens, device)
        seq_len = uncond.shape[1]
        uncond = uncond.repeat(1, num_videos_per_prompt, 1).view(batch_size * num_videos_per_prompt, seq_len, -1)
        return torch.cat([uncond, text_embeddings])

    def encode_mask(self, mask_label, num_videos_per_prompt, do_classifier_free_guidance, frame: int = 5):
        if not do_classifier_free_guidance:
            return mask_label
        seq_len = mask_label.shape[1]
        uncond = mask_label.repeat(1, num_videos_per_prompt, 1).view(frame * num_videos_per_prompt, seq_len, -1)
        return torch.cat([uncond, mask_label])

    def decode_latents(self, latents):
        f = latents.shape[2]
        lat = (latents / 0.18215).permute(0, 2, 1, 3, 4).flatten(0, 1)
        if getattr(self.vae, "batched_decode", False):
            video = self.vae.decode(lat).sample  # the B200 AutoencoderKL decodes all frames of the clip(s) in one batch
        else:
            frames = [self.vae.decode(lat[i:i + 1]).sample for i in range(lat.shape[0])]  # one frame at a time (:279-282)
            video = torch.cat(frames)
        video = video.reshape(-1, f, *video.shape[1:]).permute(0, 2, 1, 3, 4)
        video = (video / 2 + 0.5).clamp(0, 1)
        return video.cpu().float().numpy()

    def prepare_extra_step_kwargs(self, generator, eta):
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        kw = {}
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    def check_inputs(self, prompt, height, width, callback_steps):
        if not isinstance(prompt, (str, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_steps is None or not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type"
                             f" {type(callback_steps)}.")

    def prepare_latents(self, batch_size, num_channels_latents, video_length, height, width, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_channels_latents, video_length, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an"
                             f" effective batch size of {batch_size}. Make sure the batch size matches the length of"
                             " the generators.")
        if latents is None:
            if isinstance(generator, list):
                latents = torch.cat([torch.randn(shape, generator=g, device=device, dtype=dtype) for g in generator])
                latents = latents.to(device)
            else:
                latents = torch.randn(shape, generator=generator, device=device, dtype=dtype).to(device)
        else:
            if latents.shape != shape:
                raise ValueError(f"Unexpected latents shape, got {latents.shape}, expected {shape}")
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def mask2list_label(self, mask_label, encoder_hidden_states, do_classifier_free_guidance):
        """Split the text rows into 'known frame' (mask all ones) and 'to generate' (all zeros) — :350-371."""
        flat = mask_label.reshape(mask_label.size(0), -1)
        ones, zeros = (flat == 1).all(dim=1), (flat == 0).all(dim=1)
        if not bool((ones | zeros).all()):
            raise ValueError('please check mask label')
        ones, zeros = ones.cpu(), zeros.cpu()
        return encoder_hidden_states[ones], encoder_hidden_states[zeros]

    # ---- the denoise loop ------------------------------------------------------------------------------
    def _native_ok(self, latents, callback, eta) -> bool:
        return (self.use_native_loop and isinstance(self.unet, UNet3DConditionModel)
                and isinstance(self.scheduler, DDIMScheduler) and latents.is_cuda and callback is None and eta == 0.0
                and not self.scheduler.config.clip_sample)

    def denoise(self, latents, masked_label, masked_latents, ctx, num_inference_steps, guidance_scale, eta=0.0,
                extra_step_kwargs=None, callback=None, callback_steps=1):
        """Loop of ``RCDMs_pipeline.py:480-503`` from prepared tensors to final latents.
        latents (B,4,f,h,w); masked_label (2B|B,1,f,h,w); masked_latents (2B|B,4,f,h,w); ctx (2B*f|B*f, L, D)."""
        do_cfg = guidance_scale > 1.0
        self.scheduler.set_timesteps(num_inference_steps, device=latents.device)
        timesteps = self.scheduler.timesteps
        if self._native_ok(latents, callback, eta):
            return self._denoise_native(latents, masked_label, masked_latents, ctx, timesteps, guidance_scale)
        extra_step_kwargs = extra_step_kwargs if extra_step_kwargs is not None else {"eta": eta}
        latents_dtype = latents.dtype
        n_warm = len(timesteps) - num_inference_steps * self.scheduler.order
        with self.progress_bar(total=num_inference_steps) as bar:
            for i, t in enumerate(timesteps):
                x = torch.cat([latents] * 2) if do_cfg else latents
                x = self.scheduler.scale_model_input(x, t)
                x = torch.cat([x, masked_label, masked_latents], dim=1).to(dtype=latents_dtype)
                eps = self.unet(x, t, encoder_hidden_states=ctx, return_dict=False)[0]
                if do_cfg:
                    eu, ec = eps.chunk(2)
                    eps = eu + guidance_scale * (ec - eu)
                latents = self.scheduler.step(eps, t, latents, **extra_step_kwargs).prev_sample
                if i == len(timesteps) - 1 or ((i + 1) > n_warm and (i + 1) % self.scheduler.order == 0):
                    bar.update()
                    if callback is not None and i % callback_steps == 0:
                        callback(i, t, latents)
        return latents

    def _denoise_native(self, latents, masked_label, masked_latents, ctx, timesteps, guidance_scale):
        unet, sched = self.unet, self.scheduler
        clips, _, f, h, w = latents.shape
        do_cfg = guidance_scale > 1.0
        unet._ensure_bound()
        ts = [int(t) for t in timesteps.tolist()]
        coefs = [sched.step_coefficients(t) for t in ts]
        n = len(ts)
        C = _lib.C
        ts_arr = (C.c_int64 * n)(*ts)
        at_arr = (C.c_float * n)(*[c[0] for c in coefs])
        ap_arr = (C.c_float * n)(*[c[1] for c in coefs])
        # both CFG halves carry the same mask / masked latents (torch.cat([x]*2), :432,268): pass one copy
        mask1 = masked_label[:clips].contiguous()
        ml1 = masked_latents[:clips].contiguous()
        latents = latents.contiguous()
        ctx = ctx.contiguous()
        out = torch.empty_like(latents)
        dt = _lib.torch_dtype_id
        _lib.check(_lib.lib().rcdm_denoise_loop(
            unet._handle, latents.data_ptr(), dt(latents.dtype), mask1.data_ptr(), dt(mask1.dtype), ml1.data_ptr(),
            dt(ml1.dtype), ctx.data_ptr(), dt(ctx.dtype), clips, f, h, w, ctx.shape[1], ts_arr, at_arr, ap_arr, n,
            float(guidance_scale), int(self.use_cuda_graph), out.data_ptr(), _lib.current_stream_ptr()))
        unet._planned = (2 * clips if do_cfg else clips, f, h, w, ctx.shape[1])
        return out

    @torch.no_grad()
    def __call__(self, prompt, source_img, image_embeds_1, proj_embeds_0, mask_label, video_length: Optional[int],
                 height: Optional[int] = None, width: Optional[int] = None, num_inference_steps: int = 50,
                 guidance_scale: float = 7.5, negative_prompt=None, num_videos_per_prompt: Optional[int] = 1,
                 eta: float = 0.0, generator=None, latents: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "tensor", return_dict: bool = True,
                 callback: Optional[Callable[[int, int, torch.Tensor], None]] = None,
                 callback_steps: Optional[int] = 1, **kwargs):
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        self.check_inputs(prompt, height, width, callback_steps)
        batch_size = 1  # the reference processes one clip per call (:408)
        device = self._execution_device
        do_cfg = guidance_scale > 1.0
        prompt = prompt if isinstance(prompt, list) else [prompt] * batch_size
        if negative_prompt is not None:
            negative_prompt = negative_prompt if isinstance(negative_prompt, list) else [negative_prompt] * batch_size
        text_embeddings = self._encode_prompt(prompt, device, num_videos_per_prompt, do_cfg, negative_prompt)
        dtype = text_embeddings.dtype
        reps = 2 * num_videos_per_prompt if do_cfg else 1

        # reference frames -> masked-image latents (RNG draw #1), :425-432
        src = source_img.unsqueeze(0).flatten(0, 1).to(dtype=dtype, device=device)
        masked_latents = self.vae.encode(src).latent_dist.sample(generator=generator)
        masked_latents = masked_latents.reshape(1, video_length, *masked_latents.shape[1:]).permute(0, 2, 1, 3, 4)
        masked_latents = masked_latents * 0.18215
        masked_latents = torch.cat([masked_latents] * reps) if do_cfg else masked_latents

        masked_label = mask_label.squeeze().to(dtype=dtype, device=device)
        masked_label = self.encode_mask(masked_label, num_videos_per_prompt, do_cfg, frame=video_length)
        image_embeds_1 = torch.cat([image_embeds_1] * reps) if do_cfg else image_embeds_1
        proj_embeds_0 = torch.cat([proj_embeds_0] * reps) if do_cfg else proj_embeds_0

        # rich context: fused once per clip, constant over the loop (:444-450; row order quirk preserved)
        ehs_1, ehs_0 = self.mask2list_label(masked_label, text_embeddings, do_cfg)
        feature_1 = self.local_module(image_embeds_1.to(dtype=dtype, device=device), ehs_1)
        feature_0 = self.global_module(proj_embeds_0.to(dtype=dtype, device=device), ehs_0)
        ctx = torch.cat([feature_1, feature_0], dim=0)

        self.scheduler.set_timesteps(num_inference_steps, device=device)
        latents = self.prepare_latents(batch_size * num_videos_per_prompt, 4, video_length, height, width, dtype,
                                       device, generator, latents)  # RNG draw #2
        extra_step_kwargs = self.prepare_extra_step_kwargs(generator, eta)
        lh, lw = height // self.vae_scale_factor, width // self.vae_scale_factor
        masked_label = masked_label.view(-1, 1, video_length, lh, lw)  # (2,1,5,64,64) at 512x512 (:476)

        latents = self.denoise(latents, masked_label, masked_latents, ctx, num_inference_steps, guidance_scale, eta,
                               extra_step_kwargs, callback, callback_steps)

        video = self.decode_latents(latents)
        if output_type == "tensor":
            video = torch.from_numpy(video)
        if not return_dict:
            return video
        return RCDMsPipelineOutput(videos=video)
