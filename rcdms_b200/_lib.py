"""ctypes binding of librcdm_b200.so (the C ABI declared in ``include/rcdm.h``).

The library is built in-tree by ``python -m rcdms_b200.build`` (or ``__graft_entry__.build()``).
There is no CPU / PyTorch fallback: if the shared object is missing, loading fails loudly, and every
compute entry point fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
# RCDM_LIB selects an experiment build of the same library (scripts/build_variants.sh); never a different backend
LIB_PATH = os.environ.get("RCDM_LIB") or os.path.join(HERE, "_C", "librcdm_b200.so")

DT_F32, DT_F16, DT_BF16 = 0, 1, 2
GEMM_GEGLU, GEMM_GELU, GEMM_SILU, GEMM_SIMPLE = 1, 2, 4, 8
MAX_BLOCKS = 4


class UNetConfig(C.Structure):
    """Mirror of ``struct rcdm_unet_config`` (include/rcdm.h)."""
    _fields_ = [
        ("in_channels", C.c_int), ("out_channels", C.c_int), ("num_blocks", C.c_int),
        ("block_out_channels", C.c_int * MAX_BLOCKS), ("down_has_attn", C.c_int * MAX_BLOCKS),
        ("up_has_attn", C.c_int * MAX_BLOCKS), ("layers_per_block", C.c_int), ("attention_heads", C.c_int),
        ("cross_attention_dim", C.c_int), ("norm_num_groups", C.c_int), ("norm_eps", C.c_float),
        ("flip_sin_to_cos", C.c_int), ("freq_shift", C.c_float), ("use_motion_module", C.c_int),
        ("motion_down", C.c_int * MAX_BLOCKS), ("motion_up", C.c_int * MAX_BLOCKS), ("motion_mid", C.c_int),
        ("motion_heads", C.c_int), ("motion_attn_blocks", C.c_int), ("motion_max_len", C.c_int),
        ("compute_dtype", C.c_int),
    ]


# name -> (restype, argtypes); every symbol include/rcdm.h declares
_P, _I, _F, _D, _I64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int64
SIGNATURES = {
    "rcdm_version": (C.c_char_p, []),
    "rcdm_last_error": (C.c_char_p, []),
    "rcdm_device_count": (_I, []),
    "rcdm_kernel_launches": (C.c_uint64, []),
    "rcdm_debug_set_option": (_I, [C.c_char_p, _I]),
    "rcdm_set_stream_k_min": (_I, [_I]),
    "rcdm_set_gemm_pair": (_I, [_I]),
    "rcdm_unet_create": (_I, [C.POINTER(UNetConfig), C.POINTER(_P)]),
    "rcdm_unet_destroy": (None, [_P]),
    "rcdm_unet_num_weights": (_I, [_P]),
    "rcdm_unet_weight_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_I64), C.POINTER(_I)]),
    "rcdm_unet_load_weight": (_I, [_P, C.c_char_p, _P, _I, C.POINTER(_I64), _I, _P]),
    "rcdm_unet_load_weights": (_I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(_P), C.POINTER(_I), C.POINTER(_I64),
                                    C.POINTER(_I), _P]),
    "rcdm_unet_weights_missing": (_I, [_P]),
    "rcdm_unet_prepare": (_I, [_P, _I, _I, _I, _I, _I]),
    "rcdm_unet_workspace_bytes": (C.c_size_t, [_P]),
    "rcdm_unet_forward": (_I, [_P, _P, _I, _P, _D, _P, _I, _P, _I, _P]),
    "rcdm_unet_profile": (_I, [_P, _P, _I, _D, _P, _I, _P, _I, _I, _I, C.POINTER(_F), C.POINTER(_D), C.POINTER(_D),
                               C.c_char_p, C.POINTER(_I), C.POINTER(_I), _P]),
    "rcdm_unet_read_tap": (_I64, [_P, C.c_char_p, _P, _I64, C.POINTER(_I), C.POINTER(_I), _P]),
    "rcdm_unet_enable_taps": (_I, [_P, _I]),
    "rcdm_unet_set_option": (_I, [_P, C.c_char_p, _I]),
    "rcdm_ddim_cfg_step": (_I, [_P, _I, _P, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P]),
    "rcdm_denoise_loop": (_I, [_P, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, C.POINTER(_I64),
                               C.POINTER(_F), C.POINTER(_F), _I, _F, _I, _P, _P]),
    "rcdm_gemm": (_I, [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "rcdm_gemm_ex": (_I, [_I, _P, _I, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "rcdm_masked_attn": (_I, [_I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "rcdm_prior_assemble": (_I, [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rcdm_unclip_cfg_step": (_I, [_I, _P, _P, _P, _P, _I, _I, _F, _P, _I, _P]),
    "rcdm_pack_geglu": (_I, [_I, _P, _P, _P, _P, _I, _I, _P]),
    "rcdm_conv3x3": (_I, [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rcdm_pack_conv3x3": (_I, [_I, _P, _P, _I, _I, _P]),
    "rcdm_upsample_conv3x3_weight_bytes": (C.c_size_t, [_I, _I]),
    "rcdm_upsample_conv3x3": (_I, [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "rcdm_conv3x3_small": (_I, [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rcdm_linear_small": (_I, [_I, _P, _P, _P, _P, _I64, _I, _I, _P]),
    "rcdm_upsample2x": (_I, [_I, _P, _P, _I, _I, _I, _I, _P]),
    "rcdm_softmax_rows": (_I, [_I, _P, _I, _I, _I, _F, _P]),
    "rcdm_gemm_rowstats": (_I, [_I, _P, _P, _P, _P, _P, _I, _I, _I, _P, C.POINTER(C.c_int), _P]),
    "rcdm_linear_ln_scratch_bytes": (C.c_size_t, [_I, _I, _I, _I]),
    "rcdm_linear_ln": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P]),
    "rcdm_fold_ln": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "rcdm_gemm_stats_parts": (_I, [_I, _I]),
    "rcdm_rowstats": (_I, [_I, _P, _P, _I, _I, _P]),
    "rcdm_gemm_ln": (_I, [_I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _I, _I, _F, _P, _P]),
    "rcdm_fold_proj": (_I, [_I, _P, _P, _P, _P, _P, _P, _I, _P]),
    "rcdm_gemm_cat": (_I, [_I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _P, _P]),
    "rcdm_ffn_geglu_scratch_bytes": (C.c_size_t, [_I]),
    "rcdm_ffn_geglu_ln": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _P, _P]),
    "rcdm_groupnorm_scratch_bytes": (C.c_size_t, [_I, _I, _I]),
    "rcdm_groupnorm": (_I, [_I, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P]),
    "rcdm_gn_acc_bytes": (C.c_size_t, [_I, _I]),
    "rcdm_gemm_gnstats": (_I, [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "rcdm_groupnorm_from_stats": (_I, [_I, _P, _I, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "rcdm_layernorm": (_I, [_I, _P, _P, _P, _P, _I, _I, _F, _P, _I, _I, _P]),
    "rcdm_flash_attn": (_I, [_I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rcdm_temporal_attn": (_I, [_I, _P, _P, _I, _I, _I, _I, _I, _P]),
}

_lib: Optional[C.CDLL] = None


class RcdmError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RcdmError(f"{LIB_PATH} is missing: build it with `python -m rcdms_b200.build` "
                            "(there is no CPU / PyTorch fallback for the denoise path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError => the .so does not match include/rcdm.h
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int) -> None:
    """Non-zero status -> RuntimeError carrying rcdm_last_error() (no error codes in the reference API)."""
    if status != 0:
        raise RcdmError(lib().rcdm_last_error().decode(errors="replace"))


def torch_dtype_id(dtype) -> int:
    import torch
    try:
        return {torch.float32: DT_F32, torch.float16: DT_F16, torch.bfloat16: DT_BF16}[dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {dtype}; expected float32, float16 or bfloat16") from None


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
