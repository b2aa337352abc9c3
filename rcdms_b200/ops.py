"""Thin torch-tensor wrappers over the single-kernel C entry points (``rcdm_gemm``, ``rcdm_conv3x3``,
``rcdm_groupnorm``, ``rcdm_layernorm``, ``rcdm_flash_attn``, ``rcdm_temporal_attn``, ``rcdm_ddim_cfg_step``).

They exist so each kernel can be parity-tested and profiled in isolation; the UNet forward does not go
through Python per op.  All tensors must be CUDA, contiguous, float16 or bfloat16 unless stated.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def _dt(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise RuntimeError("rcdms_b200.ops: CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError("rcdms_b200.ops: contiguous tensors only")
    return _lib.torch_dtype_id(t.dtype)


def _dt16(t: torch.Tensor) -> int:
    if t.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError(f"rcdms_b200.ops: tensor-core kernels take float16/bfloat16 storage, got {t.dtype}")
    return _dt(t)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, tile_n: int = 0, simple: bool = False) -> torch.Tensor:
    """out[M,N] = a[M,K] @ w[N,K]^T (+ bias fp32[N]) (+ residual[M,N])."""
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    _lib.check(_lib.lib().rcdm_gemm(_dt16(a), a.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual), out.data_ptr(),
                                    M, N, K, 0, tile_n, int(simple), _lib.current_stream_ptr()))
    return out


def geglu_linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, simple: bool = False) -> torch.Tensor:
    """diffusers GEGLU: h, g = (a @ w^T + b).chunk(2); h * gelu(g).  w [2J, K] in the reference layout."""
    M, K = a.shape
    N = w.shape[0]
    L = _lib.lib()
    wp = torch.empty_like(w)
    bp = torch.empty((N,), dtype=torch.float32, device=a.device)
    b32 = bias.float().contiguous()
    s = _lib.current_stream_ptr()
    _lib.check(L.rcdm_pack_geglu(_dt(w), w.data_ptr(), b32.data_ptr(), wp.data_ptr(), bp.data_ptr(), N, K, s))
    out = torch.empty((M, N // 2), dtype=a.dtype, device=a.device)
    _lib.check(L.rcdm_gemm(_dt(a), a.data_ptr(), wp.data_ptr(), bp.data_ptr(), None, out.data_ptr(), M, N, K, 1, 0,
                           int(simple), s))
    return out


def linear_rowstats(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                    residual: Optional[torch.Tensor] = None):
    """linear() whose epilogue also emits per-row (sum, sum of squares) partials of the rounded output.
    Returns (out [M,N], stats fp32 [parts, M, 2])."""
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    stats = torch.zeros((64, M, 2), dtype=torch.float32, device=a.device)
    parts = _lib.C.c_int(0)
    _lib.check(_lib.lib().rcdm_gemm_rowstats(_dt16(a), a.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual),
                                             out.data_ptr(), M, N, K, stats.data_ptr(), _lib.C.byref(parts),
                                             _lib.current_stream_ptr()))
    return out, stats[:parts.value]


def linear_ln(x: torch.Tensor, w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
              bias: Optional[torch.Tensor] = None, pe: Optional[torch.Tensor] = None, rows_per_frame: int = 1,
              geglu: bool = False, eps: float = 1e-5) -> torch.Tensor:
    """(LayerNorm(x) [+ pe[frame]]) @ w^T + bias with the LayerNorm folded around the tensor-core GEMM.
    geglu: w [2J, K] / bias in the reference layout, output h * gelu(g) with J columns."""
    M, K = x.shape
    N = w.shape[0]
    L = _lib.lib()
    s = _lib.current_stream_ptr()
    frames = pe.shape[0] if pe is not None else 1
    if geglu:
        wp = torch.empty_like(w)
        bp = torch.empty((N,), dtype=torch.float32, device=x.device)
        _lib.check(L.rcdm_pack_geglu(_dt(w), w.data_ptr(), bias.float().contiguous().data_ptr(), wp.data_ptr(),
                                     bp.data_ptr(), N, K, s))
        w, bias = wp, bp
    scratch = torch.empty((L.rcdm_linear_ln_scratch_bytes(M, N, K, frames),), dtype=torch.uint8, device=x.device)
    out = torch.empty((M, N // 2 if geglu else N), dtype=x.dtype, device=x.device)
    _lib.check(L.rcdm_linear_ln(_dt16(x), x.data_ptr(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(pe),
                                _ptr(bias), out.data_ptr(), M, N, K, int(geglu), frames, rows_per_frame, eps,
                                scratch.data_ptr(), s))
    return out


def ffn_geglu_ln(y: torch.Tensor, w1: torch.Tensor, bias1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                 w2: torch.Tensor, bias2: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """y + GEGLU(LayerNorm(y) @ w1^T + bias1) @ w2^T + bias2 in ONE kernel (320-channel transformer blocks only).
    y [M, 320]; w1 [2560, 320] / bias1 [2560] in the reference layout (h rows, then gate rows); w2 [320, 1280]."""
    M, C = y.shape
    if C != 320 or tuple(w1.shape) != (2560, 320) or tuple(w2.shape) != (320, 1280):
        raise ValueError("ffn_geglu_ln is built for C = 320")
    L = _lib.lib()
    scratch = torch.empty((L.rcdm_ffn_geglu_scratch_bytes(M),), dtype=torch.uint8, device=y.device)
    out = torch.empty_like(y)
    _lib.check(L.rcdm_ffn_geglu_ln(_dt16(y), y.data_ptr(), w1.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                   bias1.float().contiguous().data_ptr(), w2.data_ptr(), bias2.float().contiguous().data_ptr(),
                                   out.data_ptr(), M, eps, scratch.data_ptr(), _lib.current_stream_ptr()))
    return out


def conv3x3(x_nhwc: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
            residual: Optional[torch.Tensor] = None, stride: int = 1, simple: bool = False) -> torch.Tensor:
    """3x3 / pad 1 conv on channels-last x [n,h,w,cin]; weight in the reference layout [cout,cin,3,3]."""
    n, h, w, cin = x_nhwc.shape
    cout = weight.shape[0]
    L = _lib.lib()
    s = _lib.current_stream_ptr()
    wp = torch.empty((cout, 9 * cin), dtype=x_nhwc.dtype, device=x_nhwc.device)
    wsrc = weight.to(x_nhwc.dtype).contiguous()
    _lib.check(L.rcdm_pack_conv3x3(_dt(wsrc), wsrc.data_ptr(), wp.data_ptr(), cout, cin, s))
    out = torch.empty((n, h // stride, w // stride, cout), dtype=x_nhwc.dtype, device=x_nhwc.device)
    _lib.check(L.rcdm_conv3x3(_dt(x_nhwc), x_nhwc.data_ptr(), wp.data_ptr(), _ptr(bias), _ptr(residual),
                              out.data_ptr(), n, h, w, cin, cout, stride, int(simple), s))
    return out


def group_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, rows_per_stat: int,
               eps: float, silu: bool) -> torch.Tensor:
    """x [rows, C] tokens; statistics over blocks of rows_per_stat rows x (C/groups) channels."""
    rows, C = x.shape
    L = _lib.lib()
    scratch = torch.zeros((L.rcdm_groupnorm_scratch_bytes(rows, rows_per_stat, groups),), dtype=torch.uint8,
                          device=x.device)
    out = torch.empty_like(x)
    _lib.check(L.rcdm_groupnorm(_dt(x), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), rows, C,
                                groups, rows_per_stat, eps, int(silu), scratch.data_ptr(),
                                _lib.current_stream_ptr()))
    return out


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
               pe: Optional[torch.Tensor] = None, rows_per_frame: int = 1, frames: int = 1) -> torch.Tensor:
    rows, C = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.lib().rcdm_layernorm(_dt(x), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), rows,
                                         C, eps, _ptr(pe), rows_per_frame, frames, _lib.current_stream_ptr()))
    return out


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, simple: bool = False) -> torch.Tensor:
    """q [batch, S_q, heads*d], k/v [batch, S_kv, heads*d] (separate contiguous tensors) -> [batch, S_q, heads*d]."""
    b, sq, c = q.shape
    skv = k.shape[1]
    d = c // heads
    if k.shape != v.shape:
        raise ValueError("k and v must have the same shape")
    out = torch.empty_like(q)
    # k and v are separate allocations: interleave into one [b*S_kv, 2c] buffer as the fused projection would
    kv = torch.cat([k, v], dim=-1).contiguous()
    _lib.check(_lib.lib().rcdm_flash_attn(_dt(q), q.data_ptr(), c, kv.data_ptr(), kv.data_ptr() + c * q.element_size(),
                                          2 * c, out.data_ptr(), c, b, heads, sq, skv, d, int(simple),
                                          _lib.current_stream_ptr()))
    return out


def temporal_attention(qkv: torch.Tensor, batch: int, frames: int, hw: int, heads: int) -> torch.Tensor:
    """qkv [(b f hw), 3C] -> [(b f hw), C]: attention over the frame axis at every (b, hw, head)."""
    rows, c3 = qkv.shape
    C = c3 // 3
    out = torch.empty((rows, C), dtype=qkv.dtype, device=qkv.device)
    _lib.check(_lib.lib().rcdm_temporal_attn(_dt(qkv), qkv.data_ptr(), out.data_ptr(), batch, frames, hw, heads,
                                             C // heads, _lib.current_stream_ptr()))
    return out


def ddim_cfg_step(eps: torch.Tensor, latents_f32: torch.Tensor, alpha_bar_t: float, alpha_bar_prev: float,
                  guidance_scale: float, do_cfg: bool, latents_dtype: torch.dtype,
                  mask: Optional[torch.Tensor] = None, masked_latents: Optional[torch.Tensor] = None,
                  next_dtype: Optional[torch.dtype] = None):
    """Fused CFG + DDIM step; returns (latents in latents_dtype, next 9-channel UNet input or None).
    latents_f32 (clips,4,f,h,w) is updated in place."""
    clips, _, f, h, w = latents_f32.shape
    lat_out = torch.empty(latents_f32.shape, dtype=latents_dtype, device=eps.device)
    nxt = None
    if mask is not None:
        nb = 2 * clips if do_cfg else clips
        nxt = torch.empty((nb, 9, f, h, w), dtype=next_dtype or eps.dtype, device=eps.device)
    _lib.check(_lib.lib().rcdm_ddim_cfg_step(
        eps.data_ptr(), _dt(eps), latents_f32.data_ptr(), lat_out.data_ptr(), _lib.torch_dtype_id(latents_dtype),
        _ptr(nxt), _lib.torch_dtype_id(nxt.dtype) if nxt is not None else 0, _ptr(mask),
        _dt(mask) if mask is not None else 0, _ptr(masked_latents),
        _dt(masked_latents) if masked_latents is not None else 0, clips, f, h, w, int(do_cfg), float(guidance_scale),
        float(alpha_bar_t), float(alpha_bar_prev), _lib.current_stream_ptr()))
    return lat_out, nxt


# ---- stage-1 frame prior kernels ------------------------------------------------------------
def linear_ex(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
              residual: Optional[torch.Tensor] = None, act: Optional[str] = None, simple: bool = False,
              rows: Optional[int] = None, lda: Optional[int] = None, a_offset: int = 0) -> torch.Tensor:
    """rcdm_gemm_ex: out = act(a @ w^T + bias) (+ residual); act in {None, "gelu", "silu"}.  ``rows`` / ``lda`` /
    ``a_offset`` (elements) select a strided row subset of ``a`` (e.g. the last token of every sample)."""
    K = w.shape[1]
    N = w.shape[0]
    M = a.shape[0] if rows is None else rows
    flags = {None: 0, "gelu": _lib.GEMM_GELU, "silu": _lib.GEMM_SILU}[act] | (_lib.GEMM_SIMPLE if simple else 0)
    out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    _lib.check(_lib.lib().rcdm_gemm_ex(_dt16(a), a.data_ptr() + a_offset * a.element_size(), lda or K, w.data_ptr(),
                                       _ptr(bias), _ptr(residual), 0, out.data_ptr(), 0, M, N, K, flags,
                                       _lib.current_stream_ptr()))
    return out


def fold_ln(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor] = None,
            pe: Optional[torch.Tensor] = None):
    """rcdm_fold_ln (load-time half of the folded LayerNorm): -> (wf [N, K] 16 bit, c [frames, N] fp32)."""
    N, K = w.shape
    frames = pe.shape[0] if pe is not None else 1
    wf = torch.empty_like(w)
    c = torch.empty((frames, N), dtype=torch.float32, device=w.device)
    _lib.check(_lib.lib().rcdm_fold_ln(_dt16(w), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(pe), _ptr(bias),
                                       wf.data_ptr(), c.data_ptr(), N, K, frames, _lib.current_stream_ptr()))
    return wf, c


def rowstats(x: torch.Tensor) -> torch.Tensor:
    """rcdm_rowstats: single-part (sum, sum of squares) per row, [1, M, 2] fp32."""
    M, K = x.shape
    st = torch.empty((1, M, 2), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().rcdm_rowstats(_dt16(x), x.data_ptr(), st.data_ptr(), M, K, _lib.current_stream_ptr()))
    return st


def gemm_ln(a: torch.Tensor, w: torch.Tensor, vec: Optional[torch.Tensor], residual: Optional[torch.Tensor] = None,
            act: Optional[str] = None, stats_in: Optional[torch.Tensor] = None, rows_per_frame: int = 1,
            emit_stats: bool = False, eps: float = 1e-5):
    """rcdm_gemm_ln.  stats_in [parts, M, 2]: ``w`` / ``vec`` are (wf, c) of ``fold_ln`` and the LayerNorm of ``a`` is
    applied through them; else ``vec`` is the bias.  act in {None, "gelu", "silu", "geglu"}.  emit_stats: also returns
    the row statistics [parts, M, 2] of the output.  -> out or (out, stats)."""
    M, K = a.shape
    N = w.shape[0]
    flags = {None: 0, "gelu": _lib.GEMM_GELU, "silu": _lib.GEMM_SILU, "geglu": _lib.GEMM_GEGLU}[act]
    out = torch.empty((M, N // 2 if act == "geglu" else N), dtype=a.dtype, device=a.device)
    frames = vec.shape[0] if stats_in is not None else 1
    so = None
    if emit_stats:
        so = torch.empty((int(_lib.lib().rcdm_gemm_stats_parts(M, N)), M, 2), dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().rcdm_gemm_ln(_dt16(a), a.data_ptr(), K, w.data_ptr(), _ptr(vec), _ptr(residual), out.data_ptr(),
                                       M, N, K, flags, _ptr(stats_in), stats_in.shape[0] if stats_in is not None else 0,
                                       frames, rows_per_frame, eps, _ptr(so), _lib.current_stream_ptr()))
    return (out, so) if emit_stats else out


def fold_proj(wp: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, bp: torch.Tensor):
    """rcdm_fold_proj: wp [C, C], w2 [C, 4C] -> (wf [C, 5C] = [wp | wp w2], cf [C] = wp b2 + bp)."""
    C = wp.shape[0]
    wf = torch.empty((C, 5 * C), dtype=wp.dtype, device=wp.device)
    cf = torch.empty((C,), dtype=torch.float32, device=wp.device)
    _lib.check(_lib.lib().rcdm_fold_proj(_dt16(wp), wp.data_ptr(), w2.data_ptr(), b2.data_ptr(), bp.data_ptr(),
                                         wf.data_ptr(), cf.data_ptr(), C, _lib.current_stream_ptr()))
    return wf, cf


def gemm_cat(a0: torch.Tensor, a1: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
             residual: Optional[torch.Tensor] = None, emit_stats: bool = False):
    """rcdm_gemm_cat: [a0 | a1] @ w^T + bias (+ residual) without materialising the concatenation."""
    M, K0 = a0.shape
    K1 = a1.shape[1]
    N = w.shape[0]
    out = torch.empty((M, N), dtype=a0.dtype, device=a0.device)
    so = None
    if emit_stats:
        so = torch.empty((int(_lib.lib().rcdm_gemm_stats_parts(M, N)), M, 2), dtype=torch.float32, device=a0.device)
    _lib.check(_lib.lib().rcdm_gemm_cat(_dt16(a0), a0.data_ptr(), K0, a1.data_ptr(), K1, w.data_ptr(), _ptr(bias),
                                        _ptr(residual), out.data_ptr(), M, N, _ptr(so), _lib.current_stream_ptr()))
    return (out, so) if emit_stats else out


def masked_attention(qkv: torch.Tensor, heads: int, key_bias: Optional[torch.Tensor] = None,
                     causal: bool = False) -> torch.Tensor:
    """qkv [batch, S, 3C] (q | k | v) -> [batch, S, C]: softmax(q k^T / sqrt(d) + key_bias[b, j] + causal) v."""
    b, S, c3 = qkv.shape
    C = c3 // 3
    out = torch.empty((b, S, C), dtype=qkv.dtype, device=qkv.device)
    _lib.check(_lib.lib().rcdm_masked_attn(_dt16(qkv), qkv.data_ptr(), c3, _ptr(key_bias), int(causal), out.data_ptr(),
                                           C, b, heads, S, C // heads, _lib.current_stream_ptr()))
    return out


def prior_assemble(base: torch.Tensor, temb_table: torch.Tensor, hproj: torch.Tensor, pos: torch.Tensor, t_row: int,
                   h_row: int, step: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, S, C = base.shape
    x = torch.empty_like(base)
    _lib.check(_lib.lib().rcdm_prior_assemble(_dt16(base), base.data_ptr(), temb_table.data_ptr(), hproj.data_ptr(),
                                              pos.data_ptr(), x.data_ptr(), B, S, C, t_row, h_row, hproj.shape[0],
                                              _ptr(step), _lib.current_stream_ptr()))
    return x


def unclip_cfg_step(pred: torch.Tensor, latents: torch.Tensor, noise_table: torch.Tensor, coef_table: torch.Tensor,
                    do_cfg: bool, guidance_scale: float, step: torch.Tensor, advance: bool = True) -> torch.Tensor:
    """In-place CFG + UnCLIP step on ``latents`` (frames, D); returns ``latents``."""
    _lib.check(_lib.lib().rcdm_unclip_cfg_step(_dt16(latents), pred.data_ptr(), latents.data_ptr(),
                                               noise_table.data_ptr(), coef_table.data_ptr(), latents.numel(),
                                               int(do_cfg), float(guidance_scale), step.data_ptr(), int(advance),
                                               _lib.current_stream_ptr()))
    return latents


def linear_gnstats(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor],
                   hw: int):
    """linear() whose epilogue also accumulates the GroupNorm chunk statistics of its output.
    Returns (out [M, N], acc uint8 buffer for group_norm_from_stats)."""
    M, K = a.shape
    N = w.shape[0]
    L = _lib.lib()
    out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    acc = torch.zeros((L.rcdm_gn_acc_bytes(M // hw, N),), dtype=torch.uint8, device=a.device)
    _lib.check(L.rcdm_gemm_gnstats(_dt16(a), a.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual), out.data_ptr(), M, N,
                                   K, hw, acc.data_ptr(), _lib.current_stream_ptr()))
    return out, acc


def group_norm_from_stats(x0: torch.Tensor, acc0: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int,
                          rows_per_stat: int, hw: int, eps: float, silu: bool, x1: Optional[torch.Tensor] = None,
                          acc1: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm(+SiLU) of the virtual concat [x0 | x1] from the producers' statistics (one streaming pass)."""
    rows, C0 = x0.shape
    C1 = x1.shape[1] if x1 is not None else 0
    out = torch.empty((rows, C0 + C1), dtype=x0.dtype, device=x0.device)
    _lib.check(_lib.lib().rcdm_groupnorm_from_stats(_dt16(x0), x0.data_ptr(), C0, acc0.data_ptr(), _ptr(x1), C1, _ptr(acc1),
                                                    gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), rows, rows_per_stat,
                                                    hw, groups, eps, int(silu), _lib.current_stream_ptr()))
    return out


def upsample_conv3x3(x_nhwc: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv3x3(pad 1)(nearest_upsample_2x(x)) with the upsample folded into the conv (no 4x tensor, K = 4 cin)."""
    n, h, w, cin = x_nhwc.shape
    cout = weight.shape[0]
    L, s = _lib.lib(), _lib.current_stream_ptr()
    wp = torch.empty((cout, 9 * cin), dtype=x_nhwc.dtype, device=x_nhwc.device)
    wsrc = weight.to(x_nhwc.dtype).contiguous()
    _lib.check(L.rcdm_pack_conv3x3(_dt(wsrc), wsrc.data_ptr(), wp.data_ptr(), cout, cin, s))
    wf = torch.empty((L.rcdm_upsample_conv3x3_weight_bytes(cout, cin),), dtype=torch.uint8, device=x_nhwc.device)
    out = torch.empty((n, 2 * h, 2 * w, cout), dtype=x_nhwc.dtype, device=x_nhwc.device)
    _lib.check(L.rcdm_upsample_conv3x3(_dt(x_nhwc), x_nhwc.data_ptr(), wp.data_ptr(), _ptr(bias), out.data_ptr(), n, h, w,
                                       cin, cout, wf.data_ptr(), 1, s))
    return out
