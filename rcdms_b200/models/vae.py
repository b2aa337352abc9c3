"""``AutoencoderKL`` on the B200 kernels — drop-in for the diffusers class RCDMs loads as its VAE
(``stage2_batchtest_rcdms_model.py:199``; used at ``RCDMs_pipeline.py:274-287`` decode and ``:429-431`` encode).

Host side in Python like the reference's; every operation on activations is a C-ABI call of ``librcdm_b200``:
conv3x3 / 1x1 as tcgen05 implicit GEMMs (``rcdm_conv3x3``, ``rcdm_gemm``), GroupNorm(+SiLU) (``rcdm_groupnorm``), the
single-head d = C attention of the mid block as two tensor-core GEMMs around ``rcdm_softmax_rows``, nearest upsample,
and two CUDA-core kernels for the 3- / 4- / 8-channel ends (``rcdm_conv3x3_small``, ``rcdm_linear_small``).
Activations are channels-last ``[n, h, w, c]`` in the module's 16-bit dtype; NCHW only at the boundary.  All frames of a
clip are decoded in ONE batch (the reference decodes frame by frame).  No CPU / PyTorch fallback.

State-dict names / shapes: diffusers 0.24.0 (``rcdms_b200/vae_spec.py``)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import _lib
from ..vae_spec import VAE_SD15_CONFIG, vae_state_dict_spec


class _Config(dict):
    __getattr__ = dict.__getitem__


class _Params(nn.Module):
    pass


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class DiagonalGaussianDistribution:
    """mean / logvar halves of the moments; ``sample`` = mean + std * randn(generator) (diffusers models/vae.py)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKLOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist


class AutoencoderKL(nn.Module):
    batched_decode = True  # RCDMsPipeline.decode_latents hands over all frames at once

    def __init__(self, **kwargs):
        super().__init__()
        unknown = set(kwargs) - set(VAE_SD15_CONFIG)
        if unknown:
            raise TypeError(f"unexpected config keys: {sorted(unknown)}")
        cfg = {**VAE_SD15_CONFIG, **kwargs}
        cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
        for c in cfg["block_out_channels"]:
            if c % 64 or c % cfg["norm_num_groups"]:
                raise NotImplementedError("block_out_channels must be multiples of 64 and of norm_num_groups")
        self._internal_dict = _Config(cfg)
        self._spec = vae_state_dict_spec(cfg)
        for name, shape in self._spec:
            *path, leaf = name.split(".")
            mod: nn.Module = self
            for p in path:
                if p not in mod._modules:
                    mod.add_module(p, _Params())
                mod = mod._modules[p]
            mod.register_parameter(leaf, nn.Parameter(torch.empty(shape), requires_grad=False))
        self._packed = None
        self._packed_key = None

    @property
    def config(self) -> _Config:
        return self._internal_dict

    @property
    def dtype(self) -> torch.dtype:
        return self.post_quant_conv.weight.dtype

    @property
    def device(self) -> torch.device:
        return self.post_quant_conv.weight.device

    @classmethod
    def from_config(cls, config: Dict, **kwargs) -> "AutoencoderKL":
        keep = {k: v for k, v in dict(config).items() if k in VAE_SD15_CONFIG}
        keep.update({k: v for k, v in kwargs.items() if k in VAE_SD15_CONFIG})
        return cls(**keep)

    def enable_slicing(self):  # API compatibility: the batched decode needs no slicing
        return None

    def disable_slicing(self):
        return None

    # ---- weights in kernel layout (packed once per (dtype, device, parameter versions)) ------------------------------
    def refresh_weights(self) -> None:
        """Force re-packing (needed after in-place edits through ``param.data``, which do not bump the version)."""
        self._packed_key = None

    def _prep(self):
        dt = self.dtype
        if dt not in (torch.float16, torch.bfloat16):
            raise TypeError(f"AutoencoderKL on B200 computes in float16 / bfloat16; call .half() first (dtype is {dt})")
        if self.device.type != "cuda":
            raise RuntimeError("AutoencoderKL needs the module on a CUDA device (there is no CPU fallback)")
        sd = dict(self.state_dict(keep_vars=True))
        key = (dt, str(self.device), tuple((t.data_ptr(), t._version) for t in sd.values()))
        if key == self._packed_key:
            return self._packed
        L, s = _lib.lib(), _lib.current_stream_ptr()
        did = _lib.torch_dtype_id(dt)
        P = {}
        for name, t in sd.items():
            t = t.detach()
            if name.endswith(".bias") or (t.dim() == 1):
                P[name] = t.float().contiguous()  # epilogue vectors / norm parameters are fp32
            elif t.dim() == 4 and t.shape[-1] == 3:
                cout, cin = t.shape[0], t.shape[1]
                wp = torch.empty((cout, 9 * cin), dtype=dt, device=t.device)
                src = t.to(dt).contiguous()
                _lib.check(L.rcdm_pack_conv3x3(did, src.data_ptr(), wp.data_ptr(), cout, cin, s))
                P[name] = wp
            else:  # 1x1 convs and Linear layers: [out, in] matrices
                P[name] = t.to(dt).reshape(t.shape[0], t.shape[1]).contiguous()
        # attention: the softmax scale C^-0.5 goes into rcdm_softmax_rows; V is produced transposed (W_v x^T) so that
        # P V is a plain A W^T GEMM, its bias added after P V (rows of P sum to one)
        torch.cuda.current_stream().synchronize()
        self._packed, self._packed_key = P, key
        return P

    # ---- kernels -----------------------------------------------------------------------------------------------------
    @staticmethod
    def _gn(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, groups: int, silu: bool) -> torch.Tensor:
        n, h, w, c = x.shape
        L = _lib.lib()
        rows = n * h * w
        scratch = torch.zeros((L.rcdm_groupnorm_scratch_bytes(rows, h * w, groups),), dtype=torch.uint8, device=x.device)
        out = torch.empty_like(x)
        _lib.check(L.rcdm_groupnorm(_lib.torch_dtype_id(x.dtype), x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(),
                                    rows, c, groups, h * w, 1e-6, int(silu), scratch.data_ptr(), _lib.current_stream_ptr()))
        return out

    @staticmethod
    def _conv(x: torch.Tensor, wp: torch.Tensor, b: torch.Tensor, res: Optional[torch.Tensor] = None,
              stride: int = 1) -> torch.Tensor:
        n, h, w, cin = x.shape
        cout = wp.shape[0]
        L, s = _lib.lib(), _lib.current_stream_ptr()
        did = _lib.torch_dtype_id(x.dtype)
        if cin < 64:  # conv_in: 3 / 4 input channels
            out = torch.empty((n, h, w, cout), dtype=x.dtype, device=x.device)
            _lib.check(L.rcdm_conv3x3_small(did, x.data_ptr(), wp.data_ptr(), b.data_ptr(), out.data_ptr(), n, h, w, cin,
                                            cout, s))
            return out
        so = 2 if stride != 1 else 1
        out = torch.empty((n, h // so, w // so, cout), dtype=x.dtype, device=x.device)
        _lib.check(L.rcdm_conv3x3(did, x.data_ptr(), wp.data_ptr(), b.data_ptr(), res.data_ptr() if res is not None else None,
                                  out.data_ptr(), n, h, w, cin, cout, stride, 0, s))
        return out

    @staticmethod
    def _linear(a: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], res: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        M, K = a.shape
        N = w.shape[0]
        if out is None:
            out = torch.empty((M, N), dtype=a.dtype, device=a.device)
        _lib.check(_lib.lib().rcdm_gemm(_lib.torch_dtype_id(a.dtype), a.data_ptr(), w.data_ptr(),
                                        b.data_ptr() if b is not None else None, res.data_ptr() if res is not None else None,
                                        out.data_ptr(), M, N, K, 0, 0, 0, _lib.current_stream_ptr()))
        return out

    def _resnet(self, P, p: str, x: torch.Tensor) -> torch.Tensor:
        g = self.config.norm_num_groups
        n, h, w, cin = x.shape
        y = self._conv(self._gn(x, P[p + ".norm1.weight"], P[p + ".norm1.bias"], g, True), P[p + ".conv1.weight"],
                       P[p + ".conv1.bias"])
        y = self._gn(y, P[p + ".norm2.weight"], P[p + ".norm2.bias"], g, True)
        if (p + ".conv_shortcut.weight") in P:
            cout = P[p + ".conv_shortcut.weight"].shape[0]
            sc = self._linear(x.reshape(-1, cin), P[p + ".conv_shortcut.weight"], P[p + ".conv_shortcut.bias"])
            x = sc.reshape(n, h, w, cout)
        return self._conv(y, P[p + ".conv2.weight"], P[p + ".conv2.bias"], res=x)

    def _attention(self, P, p: str, x: torch.Tensor) -> torch.Tensor:
        n, h, w, c = x.shape
        hw = h * w
        L, s, did = _lib.lib(), _lib.current_stream_ptr(), _lib.torch_dtype_id(x.dtype)
        y = self._gn(x, P[p + ".group_norm.weight"], P[p + ".group_norm.bias"], self.config.norm_num_groups, False)
        y2 = y.reshape(n * hw, c)
        q = self._linear(y2, P[p + ".to_q.weight"], P[p + ".to_q.bias"])
        k = self._linear(y2, P[p + ".to_k.weight"], P[p + ".to_k.bias"])
        o = torch.empty((n * hw, c), dtype=x.dtype, device=x.device)
        scores = torch.empty((hw, hw), dtype=x.dtype, device=x.device)
        vt = torch.empty((c, hw), dtype=x.dtype, device=x.device)
        for i in range(n):  # one image at a time: the (hw x hw) score matrix of one frame is 32 MB at 64 x 64 latents
            sl = slice(i * hw, (i + 1) * hw)
            self._linear(q[sl], k[sl], None, out=scores)                       # q k^T
            _lib.check(L.rcdm_softmax_rows(did, scores.data_ptr(), hw, hw, hw, float(c) ** -0.5, s))
            self._linear(P[p + ".to_v.weight"], y2[sl], None, out=vt)          # (W_v x^T) = V^T without its bias
            self._linear(scores, vt, P[p + ".to_v.bias"], out=o[sl])           # P V + b_v
        out = self._linear(o, P[p + ".to_out.0.weight"], P[p + ".to_out.0.bias"], res=x.reshape(n * hw, c))
        return out.reshape(n, h, w, c)

    def _mid(self, P, p: str, x: torch.Tensor) -> torch.Tensor:
        x = self._resnet(P, p + ".resnets.0", x)
        x = self._attention(P, p + ".attentions.0", x)
        return self._resnet(P, p + ".resnets.1", x)

    # ---- public API --------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """z (n, latent, h, w) -> ``DecoderOutput(sample=(n, 3, 8h, 8w))`` in z's dtype; any n (frames are batched)."""
        if z.dim() != 4 or z.shape[1] != self.config.latent_channels:
            raise ValueError(f"expected latents of shape (n, {self.config.latent_channels}, h, w), got {tuple(z.shape)}")
        P = self._prep()
        cfg, dt = self.config, self.dtype
        L, s, did = _lib.lib(), _lib.current_stream_ptr(), _lib.torch_dtype_id(dt)
        n, lat, h, w = z.shape
        x = z.to(dt).permute(0, 2, 3, 1).contiguous()  # NCHW -> channels-last (plumbing)
        y = torch.empty_like(x)
        _lib.check(L.rcdm_linear_small(did, x.data_ptr(), P["post_quant_conv.weight"].data_ptr(),
                                       P["post_quant_conv.bias"].data_ptr(), y.data_ptr(), n * h * w, lat, lat, s))
        x = self._conv(y, P["decoder.conv_in.weight"], P["decoder.conv_in.bias"])
        x = self._mid(P, "decoder.mid_block", x)
        nb = len(cfg.block_out_channels)
        for i in range(nb):
            for j in range(cfg.layers_per_block + 1):
                x = self._resnet(P, f"decoder.up_blocks.{i}.resnets.{j}", x)
            if i < nb - 1:
                nn_, hh, ww, cc = x.shape
                up = torch.empty((nn_, 2 * hh, 2 * ww, cc), dtype=dt, device=x.device)
                _lib.check(L.rcdm_upsample2x(did, x.data_ptr(), up.data_ptr(), nn_, hh, ww, cc, s))
                x = self._conv(up, P[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                               P[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"])
        x = self._gn(x, P["decoder.conv_norm_out.weight"], P["decoder.conv_norm_out.bias"], cfg.norm_num_groups, True)
        x = self._conv(x, P["decoder.conv_out.weight"], P["decoder.conv_out.bias"])
        img = x.permute(0, 3, 1, 2).contiguous().to(z.dtype)
        return DecoderOutput(img) if return_dict else (img,)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """image (n, 3, H, W) -> ``AutoencoderKLOutput(latent_dist)`` with moments (n, 2 * latent, H / 8, W / 8)."""
        if x.dim() != 4 or x.shape[1] != self.config.in_channels:
            raise ValueError(f"expected images of shape (n, {self.config.in_channels}, H, W), got {tuple(x.shape)}")
        P = self._prep()
        cfg, dt = self.config, self.dtype
        L, s, did = _lib.lib(), _lib.current_stream_ptr(), _lib.torch_dtype_id(dt)
        h = x.to(dt).permute(0, 2, 3, 1).contiguous()
        h = self._conv(h, P["encoder.conv_in.weight"], P["encoder.conv_in.bias"])
        nb = len(cfg.block_out_channels)
        for i in range(nb):
            for j in range(cfg.layers_per_block):
                h = self._resnet(P, f"encoder.down_blocks.{i}.resnets.{j}", h)
            if i < nb - 1:  # Downsample2D(padding=0): F.pad (0,1,0,1) + stride-2 conv == stride -2 of rcdm_conv3x3
                h = self._conv(h, P[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                               P[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=-2)
        h = self._mid(P, "encoder.mid_block", h)
        h = self._gn(h, P["encoder.conv_norm_out.weight"], P["encoder.conv_norm_out.bias"], cfg.norm_num_groups, True)
        h = self._conv(h, P["encoder.conv_out.weight"], P["encoder.conv_out.bias"])
        n, hh, ww, c2 = h.shape
        m = torch.empty_like(h)
        _lib.check(L.rcdm_linear_small(did, h.data_ptr(), P["quant_conv.weight"].data_ptr(), P["quant_conv.bias"].data_ptr(),
                                       m.data_ptr(), n * hh * ww, c2, c2, s))
        moments = m.permute(0, 3, 1, 2).contiguous().to(x.dtype)
        dist = DiagonalGaussianDistribution(moments)
        return AutoencoderKLOutput(dist) if return_dict else (dist,)
