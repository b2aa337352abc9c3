"""Checkpoint ingestion for the evaluation drivers.

The reference's inference scripts consume DeepSpeed model-state files:

* stage 2 — ``torch.load(".../mp_rank_00_model_states.pt")["module"]`` split by the prefixes ``seen_module.`` /
  ``unseen_module.`` / ``unet.`` into the two context-fusion modules and the UNet, each loaded strictly
  (``stage2_batchtest_rcdms_model.py:225-243``), on top of the SD-1.5 2-D weights ``from_pretrained_2d`` loaded
  non-strictly (``src/models/unet.py:491-501``);
* stage 1 — ``["module"]`` straight into the prior (``stage1_batchtest_rcdms_model.py:102-103``).

This module reproduces that split for the drop-in modules and adds an offline *flat* weight file: one JSON header +
raw tensors already cast to the kernels' 16-bit storage dtype, memory-mapped at load time (no pickle, no fp32 -> fp16
pass, half the bytes).  The kernel-side repacking (conv taps, fused q|k|v, GEGLU interleave, LayerNorm folding) stays
in ``rcdm_unet_load_weight`` on the device, where it costs a few milliseconds.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

StateDict = Dict[str, torch.Tensor]

STAGE2_PREFIXES = ("seen_module", "unseen_module", "unet")  # order of the reference's if / elif chain
_MAGIC = b"RCDMFLAT"
_NP = {torch.float16: np.uint16, torch.bfloat16: np.uint16, torch.float32: np.float32, torch.int64: np.int64}
_NAMES = {torch.float16: "f16", torch.bfloat16: "bf16", torch.float32: "f32", torch.int64: "i64"}
_DTYPES = {v: k for k, v in _NAMES.items()}


def split_stage2_module_state(model_sd: StateDict) -> Tuple[StateDict, StateDict, StateDict, List[str]]:
    """``stage2_batchtest_rcdms_model.py:231-239``: returns (seen_module, unseen_module, unet, unmatched keys).
    Like the reference, matching is ``startswith(prefix)`` and the prefix (with its dot) is removed with
    ``str.replace`` — i.e. everywhere in the key, not only at the front."""
    seen, unseen, unet, other = {}, {}, {}, []
    for k in model_sd.keys():
        if k.startswith("seen_module"):
            seen[k.replace("seen_module.", "")] = model_sd[k]
        elif k.startswith("unseen_module"):
            unseen[k.replace("unseen_module.", "")] = model_sd[k]
        elif k.startswith("unet"):
            unet[k.replace("unet.", "")] = model_sd[k]
        else:
            other.append(k)  # the reference prints these
    return seen, unseen, unet, other


def _module_dict(path_or_sd) -> StateDict:
    if isinstance(path_or_sd, (str, os.PathLike)):
        if str(path_or_sd).endswith(".rcdmflat"):
            return load_flat(str(path_or_sd))
        obj = torch.load(path_or_sd, map_location="cpu")
    else:
        obj = path_or_sd
    return obj["module"] if isinstance(obj, dict) and "module" in obj and isinstance(obj["module"], dict) else obj


def load_stage2_checkpoint(path_or_sd, unet, local_module=None, global_module=None, strict: bool = True) -> List[str]:
    """Load a stage-2 DeepSpeed model-state file (or an already loaded dict / a ``.rcdmflat`` file) into the drop-in
    UNet and the two context-fusion modules, strictly, like ``:241-243``.  Returns the unmatched keys."""
    seen, unseen, unet_sd, other = split_stage2_module_state(_module_dict(path_or_sd))
    if local_module is not None:
        local_module.load_state_dict(seen, strict=strict)
    if global_module is not None:
        global_module.load_state_dict(unseen, strict=strict)
    unet.load_state_dict(unet_sd, strict=strict)
    return other


def load_stage1_checkpoint(path_or_sd, prior, strict: bool = True) -> None:
    """``stage1_batchtest_rcdms_model.py:102-103``: ``prior.load_state_dict(torch.load(ckpt)["module"])``."""
    prior.load_state_dict(_module_dict(path_or_sd), strict=strict)


# ---- flat weight file ----------------------------------------------------------------------------------------------
def save_flat(sd: StateDict, path: str, dtype: Optional[torch.dtype] = torch.float16, align: int = 256) -> int:
    """Write ``sd`` as [magic | u64 header length | JSON header | padding | raw tensors].  Floating-point tensors are
    cast to ``dtype`` (None keeps each tensor's dtype); every tensor starts on an ``align``-byte boundary.  Returns the
    file size."""
    entries, blobs, off = [], [], 0
    for name, t in sd.items():
        t = t.detach().cpu()
        if t.is_floating_point() and dtype is not None:
            t = t.to(dtype)
        if t.dtype not in _NP:
            raise TypeError(f"{name}: unsupported dtype {t.dtype}")
        t = t.contiguous()
        raw = t.view(torch.int16).numpy().tobytes() if t.dtype in (torch.float16, torch.bfloat16) else t.numpy().tobytes()
        off = (off + align - 1) // align * align
        entries.append(dict(name=name, shape=list(t.shape), dtype=_NAMES[t.dtype], offset=off, nbytes=len(raw)))
        blobs.append((off, raw))
        off += len(raw)
    header = json.dumps(dict(version=1, align=align, tensors=entries)).encode()
    base = (len(_MAGIC) + 8 + len(header) + align - 1) // align * align
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(struct.pack("<Q", len(header)))
        f.write(header)
        for o, raw in blobs:
            f.seek(base + o)
            f.write(raw)
        size = f.tell()
    return size


def load_flat(path: str, mmap: bool = True) -> StateDict:
    """Read a file written by ``save_flat``.  With ``mmap`` the tensors alias a read-only memory map (pages are touched
    only when the weights are copied to the device)."""
    with open(path, "rb") as f:
        if f.read(len(_MAGIC)) != _MAGIC:
            raise ValueError(f"{path}: not an RCDM flat weight file")
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode())
    align = header["align"]
    base = (len(_MAGIC) + 8 + hlen + align - 1) // align * align
    buf = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    out: StateDict = {}
    for e in header["tensors"]:
        dt = _DTYPES[e["dtype"]]
        a = buf[base + e["offset"]: base + e["offset"] + e["nbytes"]].view(_NP[dt])
        if mmap:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")  # torch warns that the memory map is not writable; it is never written
                t = torch.from_numpy(a)
        else:
            t = torch.from_numpy(a)
        if dt in (torch.float16, torch.bfloat16):
            t = t.view(torch.int16).view(dt)
        out[e["name"]] = t.reshape(e["shape"])
    return out


def convert_stage2_checkpoint(src: str, dst: str, dtype: torch.dtype = torch.float16) -> int:
    """Offline: DeepSpeed ``mp_rank_00_model_states.pt`` -> flat file with the same (prefixed) keys in ``dtype``."""
    return save_flat(_module_dict(src), dst, dtype)
