"""TEST INFRASTRUCTURE ONLY (oracle) — functional fp32 restatement of the reference stage-1 frame prior.

Follows, line by line, the reference on a plain ``{name: tensor}`` state dict:

* ``MyPriorTransformer.forward``                    ``src/models/myprior_transformer.py:275-411``
* ``BasicTransformerBlock.forward`` (attention_bias=True, activation_fn="gelu", no cross attention)
                                                    ``src/models/attention.py:479-526``
* ``CrossAttention.forward/_attention`` with the additive (causal + padding) mask
                                                    ``src/models/attention.py:113-199``
* ``TemporalTransformer3DModel.forward`` prior_state=True branch (LayerNorm ``prior_norm``, video_length = 5)
                                                    ``src/models/motion_module.py:147-174``
* ``TemporalTransformerBlock`` / ``VersatileAttention`` / ``PositionalEncoding``
                                                    ``src/models/motion_module.py:234-246,294-354,249-267``
* the sampling loop of ``Seq_Inpaint_Prior_Pipeline.__call__``  ``src/pipelines/prior_pipeline.py:283-352``
* ``Timesteps`` / ``TimestepEmbedding`` / ``FeedForward`` / ``UnCLIPScheduler`` — restated diffusers 0.24.0
  (``oracle/diffusers_restated.py``; parity unpinned, see there).

Pinned against the reference module itself (imported unmodified through ``oracle/diffusers_shim``) by
``tests/golden/prior_*.pt`` — see ``oracle/make_golden.py`` and ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

from rcdms_b200.prior_spec import PRIOR_SCHEDULER_KWARGS, PRIOR_VIDEO_LENGTH, prior_dims
from .diffusers_restated import UnCLIPSchedulerRef, get_timestep_embedding
from .unet_ref import _batch_to_heads, _ff, _heads_to_batch, _ln, _temporal_attention

SD = Dict[str, torch.Tensor]


def _lin(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _masked_attention(sd: SD, p: str, x: torch.Tensor, heads: int, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """CrossAttention.forward with biased q/k/v and an additive mask (B*heads, S, S) — attention.py:113-199."""
    q = _heads_to_batch(_lin(sd, p + ".to_q", x), heads)
    k = _heads_to_batch(_lin(sd, p + ".to_k", x), heads)
    v = _heads_to_batch(_lin(sd, p + ".to_v", x), heads)
    scale = q.shape[-1] ** -0.5
    scores = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
                           q, k.transpose(-1, -2), beta=0, alpha=scale)
    if mask is not None:
        scores = scores + mask
    probs = scores.softmax(dim=-1).to(v.dtype)
    return _lin(sd, p + ".to_out.0", _batch_to_heads(torch.bmm(probs, v), heads))


def _ff_gelu(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """diffusers FeedForward(activation_fn="gelu"): Linear -> erf GELU -> Linear."""
    return _lin(sd, p + ".net.2", F.gelu(_lin(sd, p + ".net.0.proj", x)))


def _prior_motion(sd: SD, p: str, x: torch.Tensor, heads: int, n_attn: int) -> torch.Tensor:
    """TemporalTransformer3DModel.forward, prior_state=True — motion_module.py:150-153,166-174."""
    p = p + ".temporal_transformer"
    res = x
    y = _ln(sd, p + ".prior_norm", x)
    y = _lin(sd, p + ".proj_in", y)
    t = p + ".transformer_blocks.0"
    for i in range(n_attn):
        y = _temporal_attention(sd, f"{t}.attention_blocks.{i}", _ln(sd, f"{t}.norms.{i}", y), PRIOR_VIDEO_LENGTH,
                                heads) + y
    y = _ff(sd, t + ".ff", _ln(sd, t + ".ff_norm", y)) + y
    y = _lin(sd, p + ".proj_out", y)
    return y + res


def build_attention_mask(cfg: Dict, attention_mask: Optional[torch.Tensor], dtype) -> Optional[torch.Tensor]:
    """myprior_transformer.py:160-165,386-390: (B*heads, S, S) additive mask = padding (-10000 on masked text keys,
    0 on the additional tokens) + causal (-10000 above the diagonal)."""
    if attention_mask is None:
        return None
    d = prior_dims(cfg)
    S = d["seq"]
    causal = torch.full((S, S), -10000.0, device=attention_mask.device).triu_(1)[None]
    m = (1 - attention_mask.to(dtype)) * -10000.0
    m = F.pad(m, (0, cfg["additional_embeddings"]), value=0.0)
    m = (m[:, None, :] + causal).to(dtype)
    return m.repeat_interleave(d["heads"], dim=0)


def prior_forward(sd: SD, cfg: Dict, hidden_states: torch.Tensor, timestep, proj_embedding: torch.Tensor,
                  encoder_hidden_states: torch.Tensor, proj_embedding1: torch.Tensor, mask_label: torch.Tensor,
                  attention_mask: Optional[torch.Tensor] = None,
                  taps: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """MyPriorTransformer.forward — myprior_transformer.py:275-411.  Returns predicted_image_embedding (B, clip_dim)."""
    d = prior_dims(cfg)
    B = hidden_states.shape[0]
    dtype = sd["proj_in.weight"].dtype
    t = torch.as_tensor(timestep, device=hidden_states.device).reshape(-1)
    t = t * torch.ones(B, dtype=t.dtype, device=t.device)
    t_proj = get_timestep_embedding(t, d["inner"], flip_sin_to_cos=True, downscale_freq_shift=0).to(dtype)
    temb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", t_proj)))
    if "embedding_proj_norm.weight" in sd:
        proj_embedding = _ln(sd, "embedding_proj_norm", proj_embedding)
    pe0 = _lin(sd, "embedding_proj", proj_embedding)
    pe1 = _lin(sd, "embedding_proj1", proj_embedding1)
    ml = _lin(sd, "embedding_proj2", mask_label)
    ehs = _lin(sd, "encoder_hidden_states_proj", encoder_hidden_states)
    h = _lin(sd, "proj_in", hidden_states)

    def tok(x):
        return x[:, None, :] if x.dim() == 2 else x

    parts = [ehs, tok(pe0), tok(pe1), tok(ml), temb[:, None, :], tok(h)]
    if "prd_embedding" in sd:
        parts.append(sd["prd_embedding"].to(h.dtype).expand(B, -1, -1))
    x = torch.cat(parts, dim=1) + sd["positional_embedding"].to(h.dtype)
    mask = build_attention_mask(cfg, attention_mask, x.dtype)
    if "norm_in.weight" in sd:
        x = _ln(sd, "norm_in", x)
    if taps is not None:
        taps["tokens"] = x
    for i in range(d["layers"]):
        p = f"transformer_blocks.{2 * i}"
        x = _masked_attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), d["heads"], mask) + x
        x = _ff_gelu(sd, p + ".ff", _ln(sd, p + ".norm3", x)) + x
        if taps is not None:
            taps[p] = x
        if d["motion"]:
            p = f"transformer_blocks.{2 * i + 1}"
            x = _prior_motion(sd, p, x, d["motion_heads"], d["n_tattn"])
            if taps is not None:
                taps[p] = x
    x = _ln(sd, "norm_out", x)[:, -1]
    return _lin(sd, "proj_to_clip_embeddings", x)


CLIP_MEAN, CLIP_STD = -0.016, 0.415  # myprior_transformer.py:170-171 (post_process_latents :413-415)


def make_prior_scheduler() -> UnCLIPSchedulerRef:
    return UnCLIPSchedulerRef(**PRIOR_SCHEDULER_KWARGS)


def prior_loop(prior: Callable, latents: torch.Tensor, prompt_embeds: torch.Tensor, text_hidden: torch.Tensor,
               text_mask: torch.Tensor, imgs_proj_embeds1: torch.Tensor, mask_label: torch.Tensor,
               num_inference_steps: int, guidance_scale: float = 4.0, generator=None,
               noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """prior_pipeline.py:283-352 from the encoded prompt on.  ``prior(x, t, proj, ehs, proj1, mask_label, text_mask)``;
    prompt_embeds / text_hidden / text_mask already hold [negative, positive] rows when CFG is on (:225-230).
    ``noise`` (steps-1, F, D), optional: the variance noise per step (else drawn from ``generator``)."""
    sched = make_prior_scheduler()
    sched.set_timesteps(num_inference_steps)
    ts = sched.timesteps
    cfg_on = guidance_scale > 1
    p1 = torch.cat([imgs_proj_embeds1] * 2) if cfg_on else imgs_proj_embeds1     # :297
    ml = torch.cat([mask_label] * 2) if cfg_on else mask_label                   # :298
    latents = latents * sched.init_noise_sigma                                   # :111
    for i, t in enumerate(ts):
        x = torch.cat([latents] * 2) if cfg_on else latents                      # :303
        pred = prior(x, t, prompt_embeds, text_hidden, p1, ml, text_mask)        # :305-314
        if cfg_on:
            pu, pt = pred.chunk(2)                                               # :317
            pred = pu + guidance_scale * (pt - pu)                               # :318-320
        prev_t = None if i + 1 == ts.shape[0] else ts[i + 1]                     # :322-325
        vn = noise[i] if (noise is not None and int(t) > 0) else None
        latents = sched.step(pred, timestep=t, sample=latents, generator=generator, prev_timestep=prev_t,
                             variance_noise=vn).prev_sample                      # :327-333
    return latents * CLIP_STD + CLIP_MEAN                                        # :345


def prior_loop_oracle(sd: SD, cfg: Dict, **kw) -> torch.Tensor:
    return prior_loop(lambda x, t, pe, ehs, p1, ml, tm: prior_forward(sd, cfg, x, t, pe, ehs, p1, ml, tm), **kw)
