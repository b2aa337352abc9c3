"""GPU parity of the B200 AutoencoderKL (C-ABI kernels) against the fp32 oracle restatement fed the same 16-bit-rounded
weights / inputs.  Bound like the UNet: the reference's own half-precision noise floor."""
import pytest
import torch

from oracle import vae_ref
from rcdms_b200.models import AutoencoderKL
from rcdms_b200.vae_spec import synthetic_vae_state_dict, vae_full_config, vae_tiny_config

pytestmark = pytest.mark.gpu


def _model(cfg, dtype, seed=0):
    sd = synthetic_vae_state_dict(cfg, seed=seed)
    m = AutoencoderKL.from_config(cfg)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda", dtype), sd


def _floor_check(y, ref, half):
    d, fl = (y.float() - ref).abs(), (half.float() - ref).abs()
    assert torch.isfinite(y).all()
    assert d.max().item() <= max(3 * fl.max().item(), 5e-3 * ref.abs().max().item()), (d.max().item(), fl.max().item())
    assert d.mean().item() <= 2 * fl.mean().item() + 1e-4 * ref.abs().mean().item(), (d.mean().item(), fl.mean().item())


@pytest.mark.parametrize("cfg_fn,n,h,dtype", [(vae_tiny_config, 5, 16, torch.float16), (vae_tiny_config, 2, 32, torch.bfloat16),
                                               (vae_full_config, 5, 16, torch.float16), (vae_full_config, 1, 64, torch.float16)])
def test_decode_matches_oracle(cfg_fn, n, h, dtype):
    cfg = cfg_fn()
    m, sd = _model(cfg, dtype)
    g = torch.Generator().manual_seed(3)
    z = (torch.randn((n, 4, h, h), generator=g) / 0.18215 * 0.2).to("cuda", dtype)
    y = m.decode(z).sample
    up = 2 ** (len(cfg["block_out_channels"]) - 1)
    assert y.shape == (n, 3, h * up, h * up) and y.dtype == dtype
    sdr = {k: v.to(dtype).to("cuda", torch.float32) for k, v in sd.items()}
    sdh = {k: v.to("cuda", dtype) for k, v in sd.items()}
    with torch.no_grad():
        ref = vae_ref.vae_decode(sdr, cfg, z.float())
        half = vae_ref.vae_decode(sdh, cfg, z)
    _floor_check(y, ref, half)
    # batched decode == frame-by-frame decode (the reference's loop, RCDMs_pipeline.py:279-282) up to the summation order
    # of the stream-K GEMM decomposition, whose split points depend on the number of tiles in the launch
    if n > 1:
        one = m.decode(z[1:2]).sample
        assert (y[1:2].float() - one.float()).abs().max().item() <= 4e-3 * ref.abs().max().item()


@pytest.mark.parametrize("cfg_fn,n,hw,dtype", [(vae_tiny_config, 5, 64, torch.float16), (vae_full_config, 1, 256, torch.float16)])
def test_encode_matches_oracle(cfg_fn, n, hw, dtype):
    cfg = cfg_fn()
    m, sd = _model(cfg, dtype, seed=2)
    g = torch.Generator().manual_seed(4)
    x = (torch.rand((n, 3, hw, hw), generator=g) * 2 - 1).to("cuda", dtype)
    dist = m.encode(x).latent_dist
    sdr = {k: v.to(dtype).to("cuda", torch.float32) for k, v in sd.items()}
    sdh = {k: v.to("cuda", dtype) for k, v in sd.items()}
    with torch.no_grad():
        ref = vae_ref.vae_encode_moments(sdr, cfg, x.float())
        half = vae_ref.vae_encode_moments(sdh, cfg, x)
    assert dist.parameters.shape == ref.shape
    _floor_check(dist.parameters, ref, half)
    gen = torch.Generator(device="cuda").manual_seed(7)
    s1 = dist.sample(generator=gen)
    gen.manual_seed(7)
    noise = torch.randn(dist.mean.shape, generator=gen, device="cuda", dtype=dtype)
    assert torch.equal(s1, dist.mean + dist.std * noise)


def test_pipeline_decode_latents_batches_frames():
    from rcdms_b200.pipelines.RCDMs_pipeline import RCDMsPipeline
    cfg = vae_tiny_config()
    m, _ = _model(cfg, torch.float16)
    pipe = RCDMsPipeline.__new__(RCDMsPipeline)
    pipe.vae = m
    lat = torch.randn((1, 4, 5, 16, 16), device="cuda", dtype=torch.float16) * 0.2
    vid = pipe.decode_latents(lat)
    assert vid.shape == (1, 3, 5, 32, 32) and vid.min() >= 0.0 and vid.max() <= 1.0
    frames = torch.cat([m.decode((lat[:, :, i] / 0.18215)).sample for i in range(5)])
    ref = (frames / 2 + 0.5).clamp(0, 1).float().cpu().reshape(1, 5, 3, 32, 32).permute(0, 2, 1, 3, 4)
    assert (torch.from_numpy(vid) - ref).abs().max().item() <= 4e-3
