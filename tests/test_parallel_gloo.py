"""N > 1 host logic on CPU: world_size-2 gloo run of the clip-sharded path (partition + the single all-gather)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rcdms_b200.parallel import gather_latents, run_sharded, shard_for_rank, split_list


def test_split_list_matches_reference_partition():
    assert split_list(10, 3) == [[0, 1, 2, 3], [4, 5, 6], [7, 8, 9]]
    assert split_list(64, 8) == [list(range(8 * r, 8 * r + 8)) for r in range(8)]
    assert split_list(2, 4) == [[0], [1], [], []]
    assert sum(split_list(17, 5), []) == list(range(17))


def _fake_denoise(indices):
    # deterministic function of the CLIP index only (rank-independent), shaped like final latents
    return torch.stack([torch.full((4, 5, 2, 2), float(i)) + torch.arange(4.0).view(4, 1, 1, 1) for i in indices]) \
        if len(indices) else torch.zeros((0, 4, 5, 2, 2))


def _worker(rank, world, port, n_clips, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert shard_for_rank(n_clips) == split_list(n_clips, world)[rank]
        full = run_sharded(n_clips, _fake_denoise)
        torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [4, 5])
def test_sharded_equals_single_process(tmp_path, n_clips):
    world = 2
    port = 29500 + (os.getpid() % 500) + n_clips
    mp.spawn(_worker, args=(world, port, n_clips, str(tmp_path)), nprocs=world, join=True)
    ref = _fake_denoise(list(range(n_clips)))
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(got, ref)  # bitwise, on every rank, also with unequal shards


def test_gather_is_identity_without_process_group():
    x = torch.randn(3, 4, 5, 2, 2)
    assert gather_latents(x, 3) is x


def _fake_prior_sample(indices):
    # (clips, 5 frames, D) embeddings as a function of the clip index only, like a per-clip seeded generator would give
    out = []
    for i in indices:
        g = torch.Generator().manual_seed(42 + i)
        out.append(torch.randn((5, 16), generator=g))
    return torch.stack(out) if out else torch.zeros((0, 5, 16))


def _prior_worker(rank, world, port, n_clips, out_dir):
    from rcdms_b200.parallel import run_prior_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.save(run_prior_sharded(n_clips, _fake_prior_sample), os.path.join(out_dir, f"p{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_prior_sharded_equals_single_process(tmp_path):
    world, n_clips = 2, 3
    port = 29500 + (os.getpid() % 500) + 17
    mp.spawn(_prior_worker, args=(world, port, n_clips, str(tmp_path)), nprocs=world, join=True)
    ref = _fake_prior_sample(list(range(n_clips)))
    for r in range(world):
        assert torch.equal(torch.load(os.path.join(str(tmp_path), f"p{r}.pt")), ref)
