"""Clip-sharded multi-GPU inference.

The reference shards test clips across GPUs with zero communication: ``split_list(n_clips, n_gpus)`` gives each
spawned process a contiguous index range (``stage2_batchtest_rcdms_model.py:58-70,457-468``).  Frames of one
clip are coupled in every layer (temporal attention, cross-frame GroupNorm), so the clip is the indivisible
unit; nothing is exchanged during denoising.  This module keeps that partition and adds the single collective
the B200 build uses: one all-gather of the final latents (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def split_list(n_items: int, n_parts: int) -> List[List[int]]:
    """Contiguous split; the first ``n_items % n_parts`` parts get one extra item
    (same partition as the reference's ``split_list``)."""
    base, extra = divmod(n_items, n_parts)
    out, start = [], 0
    for r in range(n_parts):
        size = base + (1 if r < extra else 0)
        out.append(list(range(start, start + size)))
        start += size
    return out


def shard_for_rank(n_clips: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    return split_list(n_clips, world_size)[rank]


def gather_latents(local: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """All-gather per-rank final latents ``(clips_on_rank, 4, f, h, w)`` into ``(n_clips, 4, f, h, w)`` in clip order.
    Shards may be unequal (by at most one clip): every rank pads to the largest shard, one
    ``all_gather_into_tensor`` moves the data, the padding is dropped afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [len(s) for s in split_list(n_clips, world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad.contiguous(), group=group)
    parts = [buf[r * mx: r * mx + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def run_sharded(n_clips: int, denoise_clips: Callable[[Sequence[int]], torch.Tensor], group=None) -> torch.Tensor:
    """Each rank denoises its contiguous shard with ``denoise_clips(indices) -> (len(indices), 4, f, h, w)`` (inputs
    and RNG must be functions of the CLIP index, not of the rank, so the result is independent of the world size),
    then the final latents are gathered on every rank."""
    mine = shard_for_rank(n_clips)
    local = denoise_clips(mine)
    if local.shape[0] != len(mine):
        raise ValueError(f"denoise_clips returned {local.shape[0]} clips for {len(mine)} indices")
    return gather_latents(local, n_clips, group)


def run_prior_sharded(n_clips: int, sample_clips: Callable[[Sequence[int]], torch.Tensor], group=None) -> torch.Tensor:
    """Stage-1 prior, clip-sharded exactly like stage 2 (the reference spawns one process per GPU over a contiguous
    clip range, ``stage1_batchtest_rcdms_model.py`` main): ``sample_clips(indices) -> (len(indices), frames, D)`` image
    embeddings (``Seq_Inpaint_Prior_Pipeline`` per clip; inputs / generator seeded per CLIP index), then one all-gather.
    ``gather_latents`` is shape-agnostic beyond the leading clip dimension."""
    return run_sharded(n_clips, sample_clips, group)
