"""Structure of the stage-1 frame-prior transformer (reference: ``src/models/myprior_transformer.py:37-172``).

Pure-Python description shared by the host-side module mirror (``models/myprior_transformer.py``), the oracle and
the tests: the reference's configuration keys and the exact list of state-dict entries (names, shapes, order)
``MyPriorTransformer`` exposes.  No torch / CUDA dependency.

SURVEY.md §8(f) rank 1 ("next" row): the frame-prior diffusion loop directly upstream of the stage-2 denoise path
(its 1280-wide outputs are stage 2's ``proj_embeds_0``, ``stage2_batchtest_rcdms_model.py:291-294``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

from .unet_spec import RCDMS_UNET_ADDITIONAL_KWARGS

Entry = Tuple[str, Tuple[int, ...]]

# ``MyPriorTransformer.__init__`` defaults (``myprior_transformer.py:77-100``)
PRIOR_DEFAULT_CONFIG: Dict = dict(
    num_attention_heads=32, attention_head_dim=64, num_layers=20, embedding_dim=768, num_embeddings=77,
    additional_embeddings=4, dropout=0.0, time_embed_act_fn="silu", norm_in_type=None, embedding_proj_norm_type=None,
    encoder_hid_proj_type="linear", added_emb_type="prd", time_embed_dim=None, embedding_proj_dim=None,
    clip_embed_dim=None, unet_use_cross_frame_attention=None, unet_use_temporal_attention=None,
    use_motion_module=None, motion_module_type=None, motion_module_kwargs=None,
)

# kandinsky-2-2-prior ``prior/config.json`` (the checkpoint ``stage1_batchtest_rcdms_model.py:99,276`` loads; the file is
# not in the container — values from the public model card) + the two overrides ``from_pretrained_2d`` hard-codes
# (``myprior_transformer.py:428-429``: num_embeddings=91, additional_embeddings=6)
KANDINSKY22_PRIOR_CONFIG: Dict = dict(
    num_attention_heads=32, attention_head_dim=64, num_layers=20, embedding_dim=1280, num_embeddings=91,
    additional_embeddings=6, dropout=0.0, time_embed_act_fn="silu", norm_in_type=None, embedding_proj_norm_type=None,
    encoder_hid_proj_type="linear", added_emb_type="prd", time_embed_dim=None, embedding_proj_dim=None,
    clip_embed_dim=None,
)

# ``UnCLIPScheduler`` config of kandinsky-2-2-prior ``scheduler/scheduler_config.json``
# (``stage1_batchtest_rcdms_model.py:101``)
PRIOR_SCHEDULER_KWARGS: Dict = dict(num_train_timesteps=1000, variance_type="fixed_small_log", clip_sample=True,
                                    clip_sample_range=5.0, prediction_type="sample",
                                    beta_schedule="squaredcos_cap_v2")

# ``encoder_hidden_states_proj1`` input width is hard-coded (``myprior_transformer.py:131``); the layer is never used
# in ``forward`` but is part of the state dict.
_VIS_HIDDEN = 1664
PRIOR_VIDEO_LENGTH = 5  # hard-coded in the prior-state motion module (``motion_module.py:151``)


def _motion_kwargs() -> Dict:
    kw = {k: v for k, v in RCDMS_UNET_ADDITIONAL_KWARGS.items()
          if k in ("use_motion_module", "motion_module_type", "motion_module_kwargs", "unet_use_cross_frame_attention",
                   "unet_use_temporal_attention")}
    kw["motion_module_kwargs"] = dict(kw["motion_module_kwargs"])
    return kw


def prior_full_config(**overrides) -> Dict:
    """The shipped stage-1 configuration (kandinsky-2-2 prior + ``configs/testing.yaml`` motion kwargs)."""
    cfg = {**PRIOR_DEFAULT_CONFIG, **KANDINSKY22_PRIOR_CONFIG, **_motion_kwargs()}
    cfg.update(overrides)
    return cfg


def prior_tiny_config(**overrides) -> Dict:
    """Structurally identical, narrow: inner width 128 (2 heads x 64; motion heads 8 x 16), 2 layers, 11 + 6 tokens."""
    cfg = prior_full_config(num_attention_heads=2, attention_head_dim=64, num_layers=2, embedding_dim=64,
                            num_embeddings=11, additional_embeddings=6)
    cfg.update(overrides)
    return cfg


def prior_dims(cfg: Dict) -> Dict:
    inner = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    mm = cfg.get("motion_module_kwargs") or {}
    return dict(
        inner=inner, heads=cfg["num_attention_heads"], head_dim=cfg["attention_head_dim"],
        emb=cfg["embedding_dim"], proj_dim=cfg["embedding_proj_dim"] or cfg["embedding_dim"],
        clip_dim=cfg["clip_embed_dim"] or cfg["embedding_dim"], time_dim=cfg["time_embed_dim"] or inner,
        seq=cfg["num_embeddings"] + cfg["additional_embeddings"], text_len=cfg["num_embeddings"],
        layers=cfg["num_layers"], motion=bool(cfg.get("use_motion_module")),
        motion_heads=mm.get("num_attention_heads", 8),
        n_tattn=len(mm.get("attention_block_types", ())),
        max_len=mm.get("temporal_position_encoding_max_len", 24),
    )


def _lin(p: str, n: int, k: int, bias: bool = True) -> List[Entry]:
    return [(p + ".weight", (n, k))] + ([(p + ".bias", (n,))] if bias else [])


def _norm(p: str, c: int) -> List[Entry]:
    return [(p + ".weight", (c,)), (p + ".bias", (c,))]


def prior_state_dict_spec(cfg: Dict) -> List[Entry]:
    """Names / shapes in the order ``MyPriorTransformer.state_dict()`` yields them (module registration order,
    ``myprior_transformer.py:112-172``; block internals ``attention.py:390-448``, ``motion_module.py:110-145,
    204-232,270-292``)."""
    d = prior_dims(cfg)
    C, S = d["inner"], d["seq"]
    out: List[Entry] = [("positional_embedding", (1, S, C))]
    if cfg["added_emb_type"] == "prd":
        out.append(("prd_embedding", (1, 1, C)))
    out += _lin("time_embedding.linear_1", d["time_dim"], C) + _lin("time_embedding.linear_2", C, d["time_dim"])
    out += _lin("proj_in", C, d["emb"])
    if cfg["embedding_proj_norm_type"] == "layer":
        out += _norm("embedding_proj_norm", d["proj_dim"])
    out += _lin("embedding_proj", C, d["proj_dim"]) + _lin("embedding_proj1", C, d["proj_dim"])
    out += _lin("embedding_proj2", C, d["proj_dim"])
    if cfg["encoder_hid_proj_type"] == "linear":
        out += _lin("encoder_hidden_states_proj", C, d["emb"]) + _lin("encoder_hidden_states_proj1", C, _VIS_HIDDEN)
    for i in range(d["layers"]):
        p = f"transformer_blocks.{2 * i}"
        for n in ("to_q", "to_k", "to_v"):
            out += _lin(f"{p}.attn1.{n}", C, C)
        out += _lin(f"{p}.attn1.to_out.0", C, C) + _norm(f"{p}.norm1", C)
        out += _lin(f"{p}.ff.net.0.proj", 4 * C, C) + _lin(f"{p}.ff.net.2", C, 4 * C) + _norm(f"{p}.norm3", C)
        if d["motion"]:
            m = f"transformer_blocks.{2 * i + 1}.temporal_transformer"
            out += _norm(m + ".norm", C) + _norm(m + ".prior_norm", C) + _lin(m + ".proj_in", C, C)
            t = m + ".transformer_blocks.0"
            for a in range(d["n_tattn"]):
                q = f"{t}.attention_blocks.{a}"
                for n in ("to_q", "to_k", "to_v"):
                    out += _lin(f"{q}.{n}", C, C, bias=False)
                out += _lin(f"{q}.to_out.0", C, C)
                out.append((f"{q}.pos_encoder.pe", (1, d["max_len"], C)))
            for a in range(d["n_tattn"]):
                out += _norm(f"{t}.norms.{a}", C)
            out += _lin(f"{t}.ff.net.0.proj", 8 * C, C) + _lin(f"{t}.ff.net.2", C, 4 * C) + _norm(f"{t}.ff_norm", C)
            out += _lin(m + ".proj_out", C, C)
    if cfg["norm_in_type"] == "layer":
        out += _norm("norm_in", C)
    out += _norm("norm_out", C) + _lin("proj_to_clip_embeddings", d["clip_dim"], C)
    return out


PRIOR_BUFFER_SUFFIX = ".pos_encoder.pe"
# plain nn.Parameter leaves that are neither weight nor bias
PRIOR_EMBEDDING_PARAMS = ("positional_embedding", "prd_embedding")
