"""CPU tier for the VAE: the diffusers-0.24 state-dict surface, the oracle restatement's shapes and
algebra (parity unpinned: diffusers is absent; see oracle/vae_ref.py), and the drop-in module's host behaviour."""
import pytest
import torch

from oracle import vae_ref
from rcdms_b200.models import AutoencoderKL
from rcdms_b200.vae_spec import synthetic_vae_state_dict, vae_full_config, vae_state_dict_spec, vae_tiny_config


def test_state_dict_surface_of_the_sd15_vae():
    spec = vae_state_dict_spec(vae_full_config())
    names = [n for n, _ in spec]
    assert len(spec) == 248 and len(set(names)) == 248
    assert abs(sum(torch.Size(s).numel() for _, s in spec) / 1e6 - 83.65) < 0.01  # the 83.65 M parameters of the SD VAE
    d = dict(spec)
    assert d["encoder.conv_in.weight"] == (128, 3, 3, 3) and d["decoder.conv_out.weight"] == (3, 128, 3, 3)
    assert d["encoder.down_blocks.1.resnets.0.conv_shortcut.weight"] == (256, 128, 1, 1)
    assert d["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"] == (256, 512, 1, 1)
    assert d["decoder.mid_block.attentions.0.to_q.weight"] == (512, 512)
    assert d["quant_conv.weight"] == (8, 8, 1, 1) and d["post_quant_conv.weight"] == (4, 4, 1, 1)
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in d and "decoder.up_blocks.3.upsamplers.0.conv.weight" not in d
    m = AutoencoderKL.from_config(vae_full_config())
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(n, tuple(s)) for n, s in spec]


def test_oracle_shapes_and_algebra():
    cfg = vae_tiny_config()
    sd = synthetic_vae_state_dict(cfg, seed=1)
    g = torch.Generator().manual_seed(0)
    z = torch.randn((2, 4, 8, 8), generator=g)
    img = vae_ref.vae_decode(sd, cfg, z)
    assert img.shape == (2, 3, 16, 16) and torch.isfinite(img).all()  # 2 levels -> 2x up
    x = torch.randn((2, 3, 16, 16), generator=g)
    mom = vae_ref.vae_encode_moments(sd, cfg, x)
    assert mom.shape == (2, 8, 8, 8)
    # Downsample2D(padding=0): pad right / bottom only - shifting the image by one pixel must change the moments
    assert not torch.allclose(mom, vae_ref.vae_encode_moments(sd, cfg, torch.roll(x, 1, dims=-1)))
    # frames are independent (the pipeline decodes frame by frame; the B200 module batches them)
    assert torch.allclose(img[1:], vae_ref.vae_decode(sd, cfg, z[1:]), atol=1e-5)
    # DiagonalGaussianDistribution: mean + exp(0.5 clamp(logvar)) * noise
    noise = torch.randn((2, 4, 8, 8), generator=g)
    mean, logvar = mom.chunk(2, dim=1)
    assert torch.allclose(vae_ref.gaussian_sample(mom, noise=noise), mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * noise)


def test_module_has_no_cpu_fallback():
    cfg = vae_tiny_config()
    m = AutoencoderKL.from_config(cfg)
    m.load_state_dict(synthetic_vae_state_dict(cfg), strict=True)
    with pytest.raises(TypeError):
        m.decode(torch.zeros((1, 4, 8, 8)))          # fp32 module: asked to call .half() first
    with pytest.raises((RuntimeError, TypeError)):
        m.half().decode(torch.zeros((1, 4, 8, 8)))   # CPU module
    with pytest.raises(ValueError):
        m.decode(torch.zeros((1, 5, 8, 8)))
    with pytest.raises(RuntimeError):
        m.load_state_dict({"bogus": torch.zeros(1)}, strict=True)
