"""Structure of the rich-contextual UNet (reference: ``src/models/unet.py:37-251``).

Pure-Python description shared by the host-side module mirror (``models/unet.py``) and the
tests: the reference's configuration keys and the exact list of state-dict entries
(names, shapes, order) that ``UNet3DConditionModel`` exposes — 1 286 entries at the shipped
configuration.  No torch / CUDA dependency.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

# SD-1.5 ``unet/config.json`` values + the overrides ``from_pretrained_2d`` applies
# (``src/models/unet.py:476-489``) — the only configuration the reference ever builds.
SD15_UNET_CONFIG: Dict = dict(
    sample_size=64, in_channels=9, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    mid_block_type="UNetMidBlock3DCrossAttn",
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    downsample_padding=1, mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5,
    cross_attention_dim=768, attention_head_dim=8, dual_cross_attention=False, use_linear_projection=False,
    class_embed_type=None, num_class_embeds=None, upcast_attention=False, resnet_time_scale_shift="default",
    use_inflated_groupnorm=False,
)

# ``configs/testing.yaml:1-15`` (unet_additional_kwargs)
RCDMS_UNET_ADDITIONAL_KWARGS: Dict = dict(
    use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), unet_use_cross_frame_attention=False,
    unet_use_temporal_attention=False, motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=5,
                              temporal_attention_dim_div=1, zero_initialize=True),
)

# ``configs/testing.yaml:18-21`` (noise_scheduler_kwargs)
RCDMS_SCHEDULER_KWARGS: Dict = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear")

DEFAULT_CONFIG: Dict = {
    **SD15_UNET_CONFIG,
    "use_motion_module": False, "motion_module_resolutions": (1, 2, 4, 8), "motion_module_mid_block": False,
    "motion_module_decoder_only": False, "motion_module_type": None, "motion_module_kwargs": {},
    "unet_use_cross_frame_attention": None, "unet_use_temporal_attention": None,
}


def full_config(**overrides) -> Dict:
    """The shipped stage-2 configuration (SD-1.5 + testing.yaml) with optional overrides."""
    cfg = {**DEFAULT_CONFIG, **RCDMS_UNET_ADDITIONAL_KWARGS}
    cfg.update(overrides)
    return cfg


def tiny_config(**overrides) -> Dict:
    """A structurally identical but narrow configuration for fast CPU/GPU tests
    (channels 64/128/256/256 keep every code path: skip concats, shortcuts, samplers,
    head dims 8/16/32)."""
    cfg = full_config(block_out_channels=(64, 128, 256, 256), cross_attention_dim=96)
    cfg.update(overrides)
    return cfg


Entry = Tuple[str, Tuple[int, ...]]


def _resnet(p: str, cin: int, cout: int, temb: int) -> List[Entry]:
    e = [(f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)),
         (f"{p}.conv1.weight", (cout, cin, 3, 3)), (f"{p}.conv1.bias", (cout,)),
         (f"{p}.time_emb_proj.weight", (cout, temb)), (f"{p}.time_emb_proj.bias", (cout,)),
         (f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)),
         (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,))]
    if cin != cout:
        e += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]
    return e


def _ff(p: str, c: int) -> List[Entry]:
    return [(f"{p}.net.0.proj.weight", (8 * c, c)), (f"{p}.net.0.proj.bias", (8 * c,)),
            (f"{p}.net.2.weight", (c, 4 * c)), (f"{p}.net.2.bias", (c,))]


def _attn(p: str, c: int, kv: int) -> List[Entry]:
    return [(f"{p}.to_q.weight", (c, c)), (f"{p}.to_k.weight", (c, kv)), (f"{p}.to_v.weight", (c, kv)),
            (f"{p}.to_out.0.weight", (c, c)), (f"{p}.to_out.0.bias", (c,))]


def _spatial_transformer(p: str, c: int, ctx: int) -> List[Entry]:
    t = f"{p}.transformer_blocks.0"
    e = [(f"{p}.norm.weight", (c,)), (f"{p}.norm.bias", (c,)),
         (f"{p}.proj_in.weight", (c, c, 1, 1)), (f"{p}.proj_in.bias", (c,))]
    e += _attn(f"{t}.attn1", c, c) + [(f"{t}.norm1.weight", (c,)), (f"{t}.norm1.bias", (c,))]
    e += _attn(f"{t}.attn2", c, ctx) + [(f"{t}.norm2.weight", (c,)), (f"{t}.norm2.bias", (c,))]
    e += _ff(f"{t}.ff", c) + [(f"{t}.norm3.weight", (c,)), (f"{t}.norm3.bias", (c,))]
    e += [(f"{p}.proj_out.weight", (c, c, 1, 1)), (f"{p}.proj_out.bias", (c,))]
    return e


def _motion(p: str, c: int, n_attn: int, max_len: int) -> List[Entry]:
    p = f"{p}.temporal_transformer"
    t = f"{p}.transformer_blocks.0"
    e = [(f"{p}.norm.weight", (c,)), (f"{p}.norm.bias", (c,)),
         (f"{p}.prior_norm.weight", (c,)), (f"{p}.prior_norm.bias", (c,)),
         (f"{p}.proj_in.weight", (c, c)), (f"{p}.proj_in.bias", (c,))]
    for i in range(n_attn):
        e += _attn(f"{t}.attention_blocks.{i}", c, c) + [(f"{t}.attention_blocks.{i}.pos_encoder.pe", (1, max_len, c))]
    for i in range(n_attn):
        e += [(f"{t}.norms.{i}.weight", (c,)), (f"{t}.norms.{i}.bias", (c,))]
    e += _ff(f"{t}.ff", c) + [(f"{t}.ff_norm.weight", (c,)), (f"{t}.ff_norm.bias", (c,))]
    e += [(f"{p}.proj_out.weight", (c, c)), (f"{p}.proj_out.bias", (c,))]
    return e


def block_plan(cfg: Dict) -> Dict:
    """Resolve the per-block channel plan exactly as the constructor does
    (``src/models/unet.py:128-241``).  Returns a dict consumed by the spec generator, the
    oracle and (as a flat C struct) the native library."""
    boc = list(cfg["block_out_channels"])
    n = len(boc)
    lpb = cfg["layers_per_block"]
    mm = cfg.get("motion_module_kwargs") or {}
    use_mm = bool(cfg.get("use_motion_module"))
    res_ok = set(cfg.get("motion_module_resolutions", (1, 2, 4, 8)))
    down = []
    out_c = boc[0]
    for i, typ in enumerate(cfg["down_block_types"]):
        in_c, out_c = out_c, boc[i]
        down.append(dict(type=typ, cin=in_c, cout=out_c, layers=lpb, attn=typ.startswith("CrossAttn"),
                         sampler=i != n - 1,
                         motion=use_mm and (2 ** i in res_ok) and not cfg.get("motion_module_decoder_only", False)))
    up = []
    rev = boc[::-1]
    out_c = rev[0]
    for i, typ in enumerate(cfg["up_block_types"]):
        prev, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, n - 1)]
        layers = []
        for j in range(lpb + 1):
            skip = in_c if j == lpb else out_c
            rin = prev if j == 0 else out_c
            layers.append((rin, skip))
        up.append(dict(type=typ, cout=out_c, layers=layers, attn=typ.startswith("CrossAttn"), sampler=i != n - 1,
                       motion=use_mm and (2 ** (3 - i) in res_ok)))
    return dict(down=down, up=up, mid_c=boc[-1], temb=boc[0] * 4,
                mid_motion=use_mm and bool(cfg.get("motion_module_mid_block", False)),
                n_tattn=len(mm.get("attention_block_types", ())), max_len=mm.get("temporal_position_encoding_max_len", 24),
                motion_heads=mm.get("num_attention_heads", 8), ctx=cfg["cross_attention_dim"],
                heads=cfg["attention_head_dim"], groups=cfg["norm_num_groups"])


def state_dict_spec(cfg: Dict) -> List[Entry]:
    """(name, shape) for every state-dict entry, in the reference's registration order."""
    pl = block_plan(cfg)
    temb, ctx = pl["temb"], pl["ctx"]
    c0 = cfg["block_out_channels"][0]
    e: List[Entry] = [("conv_in.weight", (c0, cfg["in_channels"], 3, 3)), ("conv_in.bias", (c0,)),
                      ("time_embedding.linear_1.weight", (temb, c0)), ("time_embedding.linear_1.bias", (temb,)),
                      ("time_embedding.linear_2.weight", (temb, temb)), ("time_embedding.linear_2.bias", (temb,))]

    def motion(p, c):
        return _motion(p, c, pl["n_tattn"], pl["max_len"])

    for i, b in enumerate(pl["down"]):
        p = f"down_blocks.{i}"
        if b["attn"]:
            for j in range(b["layers"]):
                e += _spatial_transformer(f"{p}.attentions.{j}", b["cout"], ctx)
        for j in range(b["layers"]):
            e += _resnet(f"{p}.resnets.{j}", b["cin"] if j == 0 else b["cout"], b["cout"], temb)
        if b["motion"]:
            for j in range(b["layers"]):
                e += motion(f"{p}.motion_modules.{j}", b["cout"])
        if b["sampler"]:
            e += [(f"{p}.downsamplers.0.conv.weight", (b["cout"], b["cout"], 3, 3)),
                  (f"{p}.downsamplers.0.conv.bias", (b["cout"],))]
    for i, b in enumerate(pl["up"]):
        p = f"up_blocks.{i}"
        if b["attn"]:
            for j in range(len(b["layers"])):
                e += _spatial_transformer(f"{p}.attentions.{j}", b["cout"], ctx)
        for j, (rin, skip) in enumerate(b["layers"]):
            e += _resnet(f"{p}.resnets.{j}", rin + skip, b["cout"], temb)
        if b["motion"]:
            for j in range(len(b["layers"])):
                e += motion(f"{p}.motion_modules.{j}", b["cout"])
        if b["sampler"]:
            e += [(f"{p}.upsamplers.0.conv.weight", (b["cout"], b["cout"], 3, 3)),
                  (f"{p}.upsamplers.0.conv.bias", (b["cout"],))]
    mc = pl["mid_c"]
    e += _spatial_transformer("mid_block.attentions.0", mc, ctx)
    e += _resnet("mid_block.resnets.0", mc, mc, temb) + _resnet("mid_block.resnets.1", mc, mc, temb)
    if pl["mid_motion"]:
        e += motion("mid_block.motion_modules.0", mc)
    e += [("conv_norm_out.weight", (c0,)), ("conv_norm_out.bias", (c0,)),
          ("conv_out.weight", (cfg["out_channels"], c0, 3, 3)), ("conv_out.bias", (cfg["out_channels"],))]
    return e


BUFFER_SUFFIX = ".pos_encoder.pe"  # the only non-parameter entries (persistent buffers)
