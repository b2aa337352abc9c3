"""Drop-in ``UNet3DConditionModel`` whose forward is ONE call into librcdm_b200.so.

Mirrors the reference module API (``src/models/unet.py:37-508``): same constructor keywords,
``from_config`` / ``from_pretrained_2d``, ``.config``, ``.dtype``, ``.to()``, the exact 1 286
state-dict names and shapes (``load_state_dict(strict=True)`` works on a reference checkpoint,
``stage2_batchtest_rcdms_model.py:243``) and the ``forward(sample, timestep,
encoder_hidden_states, ..., return_dict)`` signature (``unet.py:322-330``).  The parameters are
plain ``nn.Parameter`` leaves grouped in ``nn.Module`` containers; there is no PyTorch
implementation of the network here — without the CUDA library the forward raises.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from .. import _lib
from ..unet_spec import BUFFER_SUFFIX, DEFAULT_CONFIG, block_plan, state_dict_spec


@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _Params(nn.Module):
    """Container node of the parameter tree (children are registered by dotted path)."""


def _c_config(cfg: Dict, dtype: torch.dtype) -> "_lib.UNetConfig":
    pl = block_plan(cfg)
    n = len(cfg["block_out_channels"])
    if n > _lib.MAX_BLOCKS:
        raise ValueError(f"at most {_lib.MAX_BLOCKS} resolution levels are supported")
    unsupported = dict(center_input_sample=False, dual_cross_attention=False, use_linear_projection=False,
                       class_embed_type=None, num_class_embeds=None, upcast_attention=False,
                       resnet_time_scale_shift="default", use_inflated_groupnorm=False, only_cross_attention=False,
                       act_fn="silu", mid_block_type="UNetMidBlock3DCrossAttn")
    for k, v in unsupported.items():
        if cfg.get(k, v) != v:
            raise NotImplementedError(f"{k}={cfg[k]!r} is not on the RCDMs stage-2 path (only {v!r})")
    if cfg.get("unet_use_cross_frame_attention") or cfg.get("unet_use_temporal_attention"):
        raise NotImplementedError("unet_use_cross_frame_attention / unet_use_temporal_attention must be false "
                                  "(configs/testing.yaml:4-5)")
    mm = cfg.get("motion_module_kwargs") or {}
    if cfg.get("use_motion_module"):
        if cfg.get("motion_module_type") != "Vanilla":
            raise ValueError("motion_module_type must be 'Vanilla' (motion_module.py:47-50)")
        if any(t != "Temporal_Self" for t in mm.get("attention_block_types", ())):
            raise NotImplementedError("only Temporal_Self attention blocks are on the RCDMs path")
        if mm.get("num_transformer_block", 1) != 1 or mm.get("temporal_attention_dim_div", 1) != 1:
            raise NotImplementedError("num_transformer_block / temporal_attention_dim_div must be 1")
        if not mm.get("temporal_position_encoding", False):
            raise NotImplementedError("temporal_position_encoding must be true (configs/testing.yaml:12)")
    c = _lib.UNetConfig()
    c.in_channels, c.out_channels, c.num_blocks = cfg["in_channels"], cfg["out_channels"], n
    for i in range(n):
        c.block_out_channels[i] = cfg["block_out_channels"][i]
        c.down_has_attn[i] = int(pl["down"][i]["attn"])
        c.up_has_attn[i] = int(pl["up"][i]["attn"])
        c.motion_down[i] = int(pl["down"][i]["motion"])
        c.motion_up[i] = int(pl["up"][i]["motion"])
    c.layers_per_block = cfg["layers_per_block"]
    heads = cfg["attention_head_dim"]
    if not isinstance(heads, int):
        if len(set(heads)) != 1:
            raise NotImplementedError("per-level attention_head_dim")
        heads = heads[0]
    c.attention_heads = heads
    c.cross_attention_dim = cfg["cross_attention_dim"]
    c.norm_num_groups, c.norm_eps = cfg["norm_num_groups"], cfg["norm_eps"]
    c.flip_sin_to_cos, c.freq_shift = int(cfg["flip_sin_to_cos"]), float(cfg["freq_shift"])
    c.use_motion_module = int(bool(cfg.get("use_motion_module")))
    c.motion_mid = int(pl["mid_motion"])
    c.motion_heads = pl["motion_heads"]
    c.motion_attn_blocks = pl["n_tattn"]
    c.motion_max_len = pl["max_len"] if cfg.get("use_motion_module") else 5
    c.compute_dtype = _lib.torch_dtype_id(dtype)
    return c


class UNet3DConditionModel(nn.Module):
    _supports_gradient_checkpointing = False  # inference-only path

    def __init__(self, **kwargs):
        super().__init__()
        unknown = set(kwargs) - set(DEFAULT_CONFIG)
        if unknown:
            raise TypeError(f"unexpected config keys: {sorted(unknown)}")
        cfg = {**DEFAULT_CONFIG, **kwargs}
        for k in ("down_block_types", "up_block_types", "block_out_channels", "motion_module_resolutions"):
            cfg[k] = tuple(cfg[k])
        self._internal_dict = _Config(cfg)
        self.sample_size = cfg["sample_size"]
        self._spec = state_dict_spec(cfg)
        for name, shape in self._spec:
            self._register(name, shape)
        self._handle = None       # native handle (created lazily, per compute dtype)
        self._debug_options = {}  # explicit per-handle debug switches (rcdm_unet_set_option); empty = product defaults
        self._handle_dtype = None
        self._bound_versions = None
        self._planned = None

    # ---- parameter tree with the reference's names ------------------------------------------------------
    def _register(self, name: str, shape: Tuple[int, ...]) -> None:
        *path, leaf = name.split(".")
        mod: nn.Module = self
        for p in path:
            if p not in mod._modules:
                mod.add_module(p, _Params())
            mod = mod._modules[p]
        if name.endswith(BUFFER_SUFFIX):
            from ..synthetic import positional_encoding
            mod.register_buffer(leaf, positional_encoding(shape[1], shape[2]))
        else:
            mod.register_parameter(leaf, nn.Parameter(torch.empty(shape), requires_grad=False))

    # ---- diffusers-like surface ---------------------------------------------------------------------------
    @property
    def config(self) -> _Config:
        return self._internal_dict

    @property
    def dtype(self) -> torch.dtype:
        return self.conv_in.weight.dtype

    @property
    def device(self) -> torch.device:
        return self.conv_in.weight.device

    @classmethod
    def from_config(cls, config: Dict, **kwargs) -> "UNet3DConditionModel":
        keep = {k: v for k, v in dict(config).items() if k in DEFAULT_CONFIG}
        keep.update({k: v for k, v in kwargs.items() if k in DEFAULT_CONFIG})
        return cls(**keep)

    @classmethod
    def from_pretrained_2d(cls, pretrained_model_path, subfolder=None, unet_additional_kwargs=None):
        """Reference: ``src/models/unet.py:465-508`` — SD-1.5 ``unet/config.json`` + 2-D weights minus
        ``conv_in.*`` loaded non-strictly, 9 input channels and the 3-D block types forced."""
        if subfolder is not None:
            pretrained_model_path = os.path.join(pretrained_model_path, subfolder)
        config_file = os.path.join(pretrained_model_path, "config.json")
        if not os.path.isfile(config_file):
            raise RuntimeError(f"{config_file} does not exist")
        with open(config_file, "r") as f:
            config = json.load(f)
        config["in_channels"] = 9
        config["down_block_types"] = ["CrossAttnDownBlock3D"] * 3 + ["DownBlock3D"]
        config["up_block_types"] = ["UpBlock3D"] + ["CrossAttnUpBlock3D"] * 3
        model = cls.from_config(config, **(unet_additional_kwargs or {}))
        model_file = os.path.join(pretrained_model_path, "diffusion_pytorch_model.bin")
        if not os.path.isfile(model_file):
            raise RuntimeError(f"{model_file} does not exist")
        state_dict = torch.load(model_file, map_location="cpu")
        match = {k: v for k, v in state_dict.items() if not k.startswith("conv_in")}
        m, u = model.load_state_dict(match, strict=False)
        print(f"### missing keys: {len(m)}; \n### unexpected keys: {len(u)};")
        return model

    def set_attention_slice(self, slice_size):
        """Accepted for API compatibility (``unet.py:253-316``): the fused attention kernel never
        materialises the score matrix, so there is nothing to slice."""
        return None

    # ---- native binding -------------------------------------------------------------------------------------
    def _versions(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def _ensure_bound(self) -> None:
        dt = self.dtype
        if dt not in (torch.float16, torch.bfloat16):
            raise TypeError("the B200 denoise path computes in float16 or bfloat16 (tensor cores); "
                            f"call .half() / .to(torch.bfloat16) first (module dtype is {dt})")
        if self.device.type != "cuda":
            raise RuntimeError("UNet3DConditionModel.forward needs the module on a CUDA device "
                               "(no CPU fallback exists for the denoise path)")
        L = _lib.lib()
        if self._handle is None or self._handle_dtype != dt:
            self._release()
            h = _lib.C.c_void_p()
            cc = _c_config(self.config, dt)
            _lib.check(L.rcdm_unet_create(_lib.C.byref(cc), _lib.C.byref(h)))
            self._handle, self._handle_dtype, self._bound_versions, self._planned = h, dt, None, None
            for k, v in self._debug_options.items():
                _lib.check(L.rcdm_unet_set_option(h, k.encode(), int(v)))
        versions = self._versions()
        if versions != self._bound_versions:
            stream = _lib.current_stream_ptr()
            items, keep = [], []
            for name, t in self.state_dict(keep_vars=True).items():
                t = t.detach()
                if not t.is_contiguous():
                    t = t.contiguous()
                    keep.append(t)
                items.append((name, t))
            n, Cc = len(items), _lib.C
            names = (Cc.c_char_p * n)(*[k.encode() for k, _ in items])
            ptrs = (Cc.c_void_p * n)(*[t.data_ptr() for _, t in items])
            dts = (Cc.c_int * n)(*[_lib.torch_dtype_id(t.dtype) for _, t in items])
            nds = (Cc.c_int * n)(*[t.dim() for _, t in items])
            dims = (Cc.c_int64 * (4 * n))()
            for i, (_, t) in enumerate(items):
                for k, d in enumerate(t.shape):
                    dims[4 * i + k] = d
            # the whole state dict in ONE packing launch
            _lib.check(L.rcdm_unet_load_weights(self._handle, n, names, ptrs, dts, dims, nds, stream))
            torch.cuda.current_stream().synchronize()
            self._bound_versions = versions

    def refresh_weights(self) -> None:
        """Force re-packing of the device weights on the next forward.  Needed after writes that bypass autograd's version
        counter (``param.data.copy_()``, LoRA / EMA merges on ``.data``): ordinary in-place ops, ``load_state_dict`` and
        ``.to()`` are detected automatically."""
        self._bound_versions = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._bound_versions = None
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._bound_versions = None
        return out

    def set_debug_option(self, name: str, value: int) -> None:
        """Explicit debug switch of the native handle ("simple": CUDA-core reference kernels for bisecting a parity
        failure; "ln_fold"; "po_fold"; "autotune").  Never set on a product path; the library reads no environment variables."""
        self._debug_options[name] = int(value)
        if self._handle is not None:
            _lib.check(_lib.lib().rcdm_unet_set_option(self._handle, name.encode(), int(value)))
            self._planned = None

    def _release(self) -> None:
        if getattr(self, "_handle", None) is not None:
            _lib.lib().rcdm_unet_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _prepare(self, b: int, f: int, h: int, w: int, L: int) -> None:
        key = (b, f, h, w, L)
        if self._planned != key:
            _lib.check(_lib.lib().rcdm_unet_prepare(self._handle, b, f, h, w, L))
            self._planned = key

    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor, class_labels: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, return_dict: bool = True):
        if class_labels is not None or attention_mask is not None:
            raise NotImplementedError("class_labels / attention_mask are None on the RCDMs stage-2 path")
        assert sample.dim() == 5, f"Expected sample to have ndim=5 (b c f h w), but got ndim={sample.dim()}."
        b, c, f, h, w = sample.shape
        if c != self.config.in_channels:
            raise ValueError(f"expected {self.config.in_channels} input channels, got {c}")
        ctx = encoder_hidden_states
        if ctx.dim() != 3 or ctx.shape[0] != b * f or ctx.shape[2] != self.config.cross_attention_dim:
            raise ValueError(f"encoder_hidden_states must be (b*f={b * f}, L, {self.config.cross_attention_dim}), "
                             f"got {tuple(ctx.shape)}")
        self._ensure_bound()
        self._prepare(b, f, h, w, ctx.shape[1])
        sample = sample.contiguous()
        ctx = ctx.contiguous()
        out = torch.empty((b, self.config.out_channels, f, h, w), dtype=sample.dtype, device=sample.device)
        t_dev, t_host = None, 0.0
        if torch.is_tensor(timestep):
            if timestep.numel() != 1:
                if not bool((timestep == timestep.flatten()[0]).all()):
                    raise NotImplementedError("per-sample timesteps: the RCDMs loop passes one scalar t per step")
                timestep = timestep.flatten()[0]
            if timestep.is_cuda and timestep.dtype == torch.int64:
                t_dev = timestep.reshape(1)  # read on the device: no host sync (RCDMs_pipeline.py:480,488)
            else:
                t_host = float(timestep)
        else:
            t_host = float(timestep)
        _lib.check(_lib.lib().rcdm_unet_forward(
            self._handle, sample.data_ptr(), _lib.torch_dtype_id(sample.dtype),
            t_dev.data_ptr() if t_dev is not None else None, t_host,
            ctx.data_ptr(), _lib.torch_dtype_id(ctx.dtype), out.data_ptr(), _lib.torch_dtype_id(out.dtype),
            _lib.current_stream_ptr()))
        if not return_dict:
            return (out,)
        return out  # the reference returns the bare tensor here too (unet.py:462-463)

    # ---- measurement -----------------------------------------------------------------------------------
    @torch.no_grad()
    def profile(self, sample: torch.Tensor, timestep: float, encoder_hidden_states: torch.Tensor, reps: int = 3):
        """Per-op CUDA-event timing of one forward (``rcdm_unet_profile``): list of dicts
        {kind, ms, flops, bytes} in launch order."""
        b, c, f, h, w = sample.shape
        ctx = encoder_hidden_states.contiguous()
        sample = sample.contiguous()
        self._ensure_bound()
        self._prepare(b, f, h, w, ctx.shape[1])
        out = torch.empty((b, self.config.out_channels, f, h, w), dtype=sample.dtype, device=sample.device)
        C = _lib.C
        cap = 4096
        ms, fl, by = (C.c_float * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
        kinds = C.create_string_buffer(16 * cap)
        dims = (C.c_int * (3 * cap))()
        n = C.c_int()
        _lib.check(_lib.lib().rcdm_unet_profile(
            self._handle, sample.data_ptr(), _lib.torch_dtype_id(sample.dtype), float(timestep), ctx.data_ptr(),
            _lib.torch_dtype_id(ctx.dtype), out.data_ptr(), _lib.torch_dtype_id(out.dtype), reps, cap, ms, fl, by,
            kinds, dims, C.byref(n), _lib.current_stream_ptr()))
        raw = kinds.raw
        return [dict(kind=raw[16 * i:16 * i + 16].split(b"\0")[0].decode(), ms=ms[i], flops=fl[i], bytes=by[i],
                     shape=(dims[3 * i], dims[3 * i + 1], dims[3 * i + 2]))
                for i in range(n.value)]

    # ---- debugging aid ----------------------------------------------------------------------------------
    def enable_taps(self, enable: bool = True) -> None:
        self._ensure_bound()
        _lib.check(_lib.lib().rcdm_unet_enable_taps(self._handle, int(enable)))
        self._planned = None

    def read_tap(self, name: str) -> torch.Tensor:
        """fp32 channels-last tokens [(b f h w), C] of an internal activation of the last forward."""
        L = _lib.lib()
        rows, ch = _lib.C.c_int(), _lib.C.c_int()
        n = L.rcdm_unet_read_tap(self._handle, name.encode(), None, 0, _lib.C.byref(rows), _lib.C.byref(ch), None)
        if n < 0:
            _lib.check(1)
        out = torch.empty((rows.value, ch.value), dtype=torch.float32, device=self.device)
        n = L.rcdm_unet_read_tap(self._handle, name.encode(), out.data_ptr(), out.numel(), _lib.C.byref(rows),
                                 _lib.C.byref(ch), _lib.current_stream_ptr())
        if n < 0:
            _lib.check(1)
        return out
