"""GPU: the native denoise loop (rcdm_denoise_loop: CUDA graph of UNet + CFG + DDIM) through the drop-in pipeline
vs (a) the same pipeline's python loop over unet()/scheduler.step(), (b) the oracle's restated loop in fp32."""
import pytest
import torch

import unet_checks as uc
from fakes import FakeTextEncoder, FakeTokenizer, FakeVAE
from oracle.loop_ref import denoise_loop
from oracle.unet_ref import unet_forward
from rcdms_b200.pipelines.RCDMs_pipeline import RCDMsPipeline, local_feature
from rcdms_b200.schedulers import DDIMScheduler
from rcdms_b200.synthetic import synthetic_clip_inputs, synthetic_state_dict
from rcdms_b200.unet_spec import RCDMS_SCHEDULER_KWARGS, tiny_config

pytestmark = pytest.mark.gpu


def _pipe(dtype=torch.float16, cfg=None, sd=None):
    cfg = cfg or tiny_config()
    unet = uc.build_model(cfg, dtype, sd)
    torch.manual_seed(0)
    D = cfg["cross_attention_dim"]
    lm = local_feature(D, 16, D, 8).to("cuda", dtype)
    gm = local_feature(D, 12, D, 8).to("cuda", dtype)
    pipe = RCDMsPipeline(FakeVAE().to("cuda", dtype), FakeTextEncoder(7, D).to("cuda", dtype), FakeTokenizer(), unet, lm,
                         gm, DDIMScheduler(**RCDMS_SCHEDULER_KWARGS))
    return cfg, pipe


def _inputs(cfg, clips, h, dtype, ctx_len=7):
    ins = [synthetic_clip_inputs(k, h, h, ctx_len=ctx_len, ctx_dim=cfg["cross_attention_dim"]) for k in range(clips)]
    lat = torch.cat([i["latents"] for i in ins]).to("cuda", dtype)
    ml = torch.cat([i["masked_latents"] for i in ins]).to("cuda", dtype)
    mask = torch.cat([i["mask"] for i in ins]).to("cuda", dtype)
    ctx = torch.cat([i["ctx"][:5] for i in ins] + [i["ctx"][5:] for i in ins]).to("cuda", dtype)
    return ins, lat, ml, mask, ctx


@pytest.mark.parametrize("clips,steps", [(1, 10), (3, 4)])
def test_native_loop_matches_python_loop_and_oracle(clips, steps):
    dtype = torch.float16
    cfg, pipe = _pipe(dtype)
    ins, lat, ml, mask, ctx = _inputs(cfg, clips, 8, dtype)
    nat = pipe.denoise(lat, torch.cat([mask] * 2), torch.cat([ml] * 2), ctx, steps, 2.0)
    nat2 = pipe.denoise(lat, torch.cat([mask] * 2), torch.cat([ml] * 2), ctx, steps, 2.0)
    assert torch.equal(nat, nat2), "graph replay must be deterministic"
    pipe.use_cuda_graph = False
    nog = pipe.denoise(lat, torch.cat([mask] * 2), torch.cat([ml] * 2), ctx, steps, 2.0)
    assert torch.equal(nat, nog), "graph and eager launches must agree bitwise"
    pipe.use_native_loop = False
    py = pipe.denoise(lat, torch.cat([mask] * 2), torch.cat([ml] * 2), ctx, steps, 2.0)
    # same UNet kernels; the fused CFG + DDIM kernel mirrors torch's per-op fp16 rounding (elementwise.cuh), so the
    # native loop and the reference-shaped python loop must agree bit for bit
    d_py = (nat.float() - py.float()).abs().max().item()
    assert torch.equal(nat, py), f"native loop vs python loop: max abs diff {d_py}"
    sd = {k: v.half().float().cuda() for k, v in synthetic_state_dict(cfg, seed=0).items()}
    with torch.no_grad():
        refs = []
        for k in range(clips):  # clips are independent: the oracle runs them one by one
            c = torch.cat([ctx[k * 5:(k + 1) * 5], ctx[(clips + k) * 5:(clips + k + 1) * 5]]).float()
            refs.append(denoise_loop(lambda x, t, cc: unet_forward(sd, cfg, x, t, cc), lat[k:k + 1].float(),
                                     mask[k:k + 1].float(), ml[k:k + 1].float(), c, steps, 2.0))
        ref = torch.cat(refs)
    # fp16 path vs the fp32 oracle after `steps` DDIM steps: the reference's own per-op fp16 rounding is part of the
    # difference (random-weight latents grow to |x| ~ 30), so the bound is relative to the reference's magnitude
    diff = (nat.float() - ref).abs()
    assert torch.isfinite(nat).all()
    assert diff.max().item() < 1e-2 * ref.abs().max().item(), (diff.max().item(), ref.abs().max().item())
    assert diff.mean().item() < 2e-3 * ref.pow(2).mean().sqrt().item(), (diff.mean().item(), ref.pow(2).mean().sqrt().item())


def test_native_loop_at_the_benchmarked_shape():
    """BASELINE config 2's loop (full-width UNet, 64x64 latents, L = 85, fp16, CFG 2.0), shortened to 5 DDIM steps:
    graph replay == eager launches == python loop bitwise, and all of them track the fp32 oracle loop."""
    from rcdms_b200.unet_spec import full_config
    dtype, steps = torch.float16, 5
    cfg = full_config()
    sd32 = synthetic_state_dict(cfg, seed=0)
    _, pipe = _pipe(dtype, cfg, sd32)
    ins, lat, ml, mask, ctx = _inputs(cfg, 1, 64, dtype, ctx_len=85)
    m2, ml2 = torch.cat([mask] * 2), torch.cat([ml] * 2)
    nat = pipe.denoise(lat, m2, ml2, ctx, steps, 2.0)
    assert torch.equal(nat, pipe.denoise(lat, m2, ml2, ctx, steps, 2.0)), "graph replay must be deterministic"
    pipe.use_cuda_graph = False
    assert torch.equal(nat, pipe.denoise(lat, m2, ml2, ctx, steps, 2.0)), "graph and eager launches must agree bitwise"
    pipe.use_native_loop = False
    py = pipe.denoise(lat, m2, ml2, ctx, steps, 2.0)
    assert torch.equal(nat, py), f"native vs python loop: {(nat.float() - py.float()).abs().max().item()}"
    sd = {k: v.half().float().cuda() for k, v in sd32.items()}
    with torch.no_grad():
        ref = denoise_loop(lambda x, t, cc: unet_forward(sd, cfg, x, t, cc), lat.float(), mask.float(), ml.float(),
                           ctx.float(), steps, 2.0)
        # the reference's own fp16 noise after the same 5 steps (oracle in eager fp16 on this GPU)
        sdh = {k: v.half() for k, v in sd.items()}
        half = denoise_loop(lambda x, t, cc: unet_forward(sdh, cfg, x, t, cc), lat, mask, ml, ctx, steps, 2.0)
    diff, floor = (nat.float() - ref).abs(), (half.float() - ref).abs()
    assert torch.isfinite(nat).all()
    assert diff.max().item() <= max(3 * floor.max().item(), 1e-2 * ref.abs().max().item()), (diff.max().item(), floor.max().item())
    assert diff.mean().item() <= 2 * floor.mean().item() + 1e-4, (diff.mean().item(), floor.mean().item())


def test_full_call_through_pipeline():
    cfg, pipe = _pipe()
    g = torch.Generator(device="cuda").manual_seed(42)
    mask = torch.zeros((5, 1, 8, 8))
    mask[0] = 1.0
    out = pipe(prompt=[f"caption {i}" for i in range(5)], source_img=torch.randn((5, 3, 64, 64)),
               image_embeds_1=torch.randn((1, 9, 16)).cuda(), proj_embeds_0=torch.randn((4, 1, 12)).cuda(),
               mask_label=mask, video_length=5, height=64, width=64, guidance_scale=2.0, num_inference_steps=3,
               generator=g)
    v = out.videos
    assert v.shape == (1, 3, 5, 64, 64) and v.dtype == torch.float32 and torch.isfinite(v).all()


def test_guidance_off_runs_single_branch():
    cfg, pipe = _pipe()
    ins, lat, ml, mask, ctx = _inputs(cfg, 1, 8, torch.float16)
    out = pipe.denoise(lat, mask, ml, ctx[:5], 3, 1.0)
    assert out.shape == lat.shape and torch.isfinite(out).all()
