"""Host-side logic of the drop-in RCDMsPipeline on CPU (python-loop path with a stand-in UNet): call surface,
mask / context plumbing (including the reference's context row-order quirk), error behaviour, and equality with
the oracle's restated loop."""
import pytest
import torch

from fakes import FakeTextEncoder, FakeTokenizer, FakeVAE
from oracle.loop_ref import denoise_loop
from rcdms_b200.pipelines.RCDMs_pipeline import RCDMsPipeline, RCDMsPipelineOutput, local_feature
from rcdms_b200.schedulers import DDIMScheduler
from rcdms_b200.unet_spec import RCDMS_SCHEDULER_KWARGS


class TinyUNet(torch.nn.Module):
    """Deterministic stand-in with the UNet call signature (NOT the product UNet: that has no CPU path)."""

    class config:
        sample_size = 4
        in_channels = 9

    def __init__(self):
        super().__init__()
        self.mix = torch.nn.Parameter(torch.randn(4, 9) * 0.3, requires_grad=False)
        self.calls = []

    def forward(self, x, t, encoder_hidden_states=None, return_dict=True):
        self.calls.append((tuple(x.shape), int(t), tuple(encoder_hidden_states.shape)))
        bias = encoder_hidden_states.mean(dim=(1, 2)).reshape(x.shape[0], x.shape[2])  # (b, f): couples ctx rows
        y = torch.einsum("oc,bcfhw->bofhw", self.mix, x) * (1 + 0.001 * float(t)) + bias[:, None, :, None, None]
        return (y,)


def make_pipe(L=7, D=96):
    torch.manual_seed(0)
    unet = TinyUNet()
    lm = local_feature(text_dim=D, vis_dim=16, hidden_dim=D, num_heads=8)
    gm = local_feature(text_dim=D, vis_dim=12, hidden_dim=D, num_heads=8)
    lm.allow_torch_path = gm.allow_torch_path = True  # CPU tier: host logic only (the product path is CUDA-only)
    pipe = RCDMsPipeline(FakeVAE(), FakeTextEncoder(L, D), FakeTokenizer(), unet, lm, gm,
                         DDIMScheduler(**RCDMS_SCHEDULER_KWARGS))
    return pipe, unet


def call_args(h=32, w=32):
    g = torch.Generator().manual_seed(42)
    mask = torch.zeros((5, 1, h // 8, w // 8))
    mask[0] = 1.0  # 'continue' mode: frame 0 known
    return dict(prompt=[f"caption number {i}" for i in range(5)], source_img=torch.randn((5, 3, h, w), generator=g),
                image_embeds_1=torch.randn((1, 9, 16), generator=g), proj_embeds_0=torch.randn((4, 1, 12), generator=g),
                mask_label=mask, video_length=5, height=h, width=w, guidance_scale=2.0, num_inference_steps=4,
                generator=torch.Generator().manual_seed(7))


def test_scheduler_config_is_patched_like_the_reference():
    pipe, _ = make_pipe()
    assert pipe.scheduler.config.steps_offset == 1 and pipe.scheduler.config.clip_sample is False
    assert pipe.vae_scale_factor == 8


def test_call_surface_and_shapes():
    pipe, unet = make_pipe()
    out = pipe(**call_args())
    assert isinstance(out, RCDMsPipelineOutput)
    v = out.videos
    assert v.shape == (1, 3, 5, 32, 32) and v.dtype == torch.float32 and v.device.type == "cpu"
    assert float(v.min()) >= 0.0 and float(v.max()) <= 1.0
    assert len(unet.calls) == 4
    shape, t0, ctx_shape = unet.calls[0]
    assert shape == (2, 9, 5, 4, 4) and ctx_shape == (10, 7, 96)
    assert [c[1] for c in unet.calls] == [751, 501, 251, 1]  # leading spacing, steps_offset = 1
    v2 = pipe(**call_args(), return_dict=False)
    assert torch.equal(v2, v)


def test_loop_equals_oracle_restatement():
    pipe, unet = make_pipe()
    g = torch.Generator().manual_seed(3)
    lat = torch.randn((1, 4, 5, 4, 4), generator=g)
    mask = torch.zeros((1, 1, 5, 4, 4))
    mask[:, :, 0] = 1
    ml = torch.randn((1, 4, 5, 4, 4), generator=g) * 0.18215
    ctx = torch.randn((10, 7, 96), generator=g)
    ours = pipe.denoise(lat, torch.cat([mask] * 2), torch.cat([ml] * 2), ctx, 6, 2.0)
    ref = denoise_loop(lambda x, t, c: unet(x, t, encoder_hidden_states=c)[0], lat, mask, ml, ctx, 6, 2.0)
    assert torch.allclose(ours, ref, atol=1e-5, rtol=1e-5)
    # guidance <= 1 disables CFG: single batch, context of f rows
    ours1 = pipe.denoise(lat, mask, ml, ctx[:5], 3, 1.0)
    ref1 = denoise_loop(lambda x, t, c: unet(x, t, encoder_hidden_states=c)[0], lat, mask, ml, ctx[:5], 3, 1.0)
    assert torch.allclose(ours1, ref1, atol=1e-5, rtol=1e-5)


def test_context_row_order_quirk_is_preserved():
    """RCDMs_pipeline.py:444-450: ctx = cat([local(image_embeds_1, ehs_1), global(proj_embeds_0, ehs_0)]) —
    rows of known-frame prompts first (uncond then cond), then the frames to generate."""
    pipe, _ = make_pipe()
    text = torch.arange(10.0).view(10, 1, 1).expand(10, 7, 96).clone()
    mask = torch.zeros((10, 4, 4))
    mask[0] = 1
    mask[5] = 1
    e1, e0 = pipe.mask2list_label(mask, text, True)
    assert e1[:, 0, 0].tolist() == [0.0, 5.0]
    assert e0[:, 0, 0].tolist() == [1.0, 2.0, 3.0, 4.0, 6.0, 7.0, 8.0, 9.0]
    mask[3, 0, 0] = 0.5
    with pytest.raises(ValueError, match="please check mask label"):
        pipe.mask2list_label(mask, text, True)


def test_error_behaviour_matches_reference():
    pipe, _ = make_pipe()
    a = call_args()
    with pytest.raises(ValueError, match="divisible by 8"):
        pipe(**{**a, "height": 30})
    with pytest.raises(ValueError, match="`prompt` has to be of type"):
        pipe(**{**a, "prompt": 3})
    with pytest.raises(ValueError, match="callback_steps"):
        pipe(**{**a, "callback_steps": 0})
    with pytest.raises(ValueError, match="Unexpected latents shape"):
        pipe(**{**a, "latents": torch.zeros((1, 4, 5, 3, 3))})
    with pytest.raises(ValueError, match="negative_prompt"):  # wrapped to [x]*1 first, then a batch-size mismatch
        pipe(**{**a, "negative_prompt": "ugly"})
    seen = []
    pipe(**a, callback=lambda i, t, lat: seen.append((i, int(t), tuple(lat.shape))), callback_steps=2)
    assert [s[0] for s in seen] == [0, 2] and seen[0][2] == (1, 4, 5, 4, 4)


def test_src_compat_imports():
    from src.models.unet import UNet3DConditionModel
    from src.pipelines.RCDMs_pipeline import RCDMsPipeline as P2
    from rcdms_b200.models import UNet3DConditionModel as U
    assert UNet3DConditionModel is U and P2 is RCDMsPipeline
