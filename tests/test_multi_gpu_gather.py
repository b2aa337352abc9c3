"""GPU, N > 1: clip-sharded denoise over the REAL pipeline (one process per GPU, NCCL) == single-GPU result, bitwise.

Partition semantics of the reference (``stage2_batchtest_rcdms_model.py:58-70,457-468``: contiguous ``split_list``, one
spawned process per GPU, nothing exchanged while denoising) plus the one collective this build adds: an all-gather of
the final latents (``rcdms_b200.parallel``).  Inputs are functions of the CLIP index, so the result must not depend on
the world size.  Skipped when the box has fewer GPUs than the case needs (run with ``gpurun --gpus N``)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _build_pipe(dtype):
    import unet_checks as uc
    from rcdms_b200.pipelines.RCDMs_pipeline import RCDMsPipeline
    from rcdms_b200.schedulers import DDIMScheduler
    from rcdms_b200.unet_spec import RCDMS_SCHEDULER_KWARGS, tiny_config
    cfg = tiny_config()
    unet = uc.build_model(cfg, dtype)

    class _VaeCfg:
        block_out_channels = (128, 256, 512, 512)

    class _Vae:
        config = _VaeCfg()
    return cfg, RCDMsPipeline(vae=_Vae(), text_encoder=None, tokenizer=None, unet=unet, local_module=None,
                              global_module=None, scheduler=DDIMScheduler(**RCDMS_SCHEDULER_KWARGS))


def _denoise_clip(pipe, cfg, k, dtype, steps):
    from rcdms_b200.synthetic import synthetic_clip_inputs
    i = synthetic_clip_inputs(k, 16, 16, ctx_len=7, ctx_dim=cfg["cross_attention_dim"])
    d = {n: v.to("cuda", dtype) for n, v in i.items()}
    return pipe.denoise(d["latents"], torch.cat([d["mask"]] * 2), torch.cat([d["masked_latents"]] * 2), d["ctx"], steps, 2.0)


def _worker(rank, world, port, n_clips, steps, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from rcdms_b200.parallel import run_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dtype = torch.float16
        cfg, pipe = _build_pipe(dtype)

        def denoise_clips(indices):
            if not len(indices):
                return torch.zeros((0, 4, 5, 16, 16), dtype=dtype, device="cuda")
            return torch.cat([_denoise_clip(pipe, cfg, k, dtype, steps) for k in indices])
        full = run_sharded(n_clips, denoise_clips)
        torch.save(full.cpu(), os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_denoise_equals_single_gpu_bitwise(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n_clips, steps = world + 1, 3  # unequal shards: rank 0 takes two clips
    port = 29600 + (os.getpid() % 300) + world
    mp.spawn(_worker, args=(world, port, n_clips, steps, str(tmp_path)), nprocs=world, join=True)
    dtype = torch.float16
    cfg, pipe = _build_pipe(dtype)
    ref = torch.cat([_denoise_clip(pipe, cfg, k, dtype, steps) for k in range(n_clips)]).cpu()
    assert ref.shape == (n_clips, 4, 5, 16, 16) and torch.isfinite(ref).all()
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(got, ref), f"rank {r}: sharded result differs from the single-GPU result"
