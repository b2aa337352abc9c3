"""Closed-form known answers for the DDIM schedule / Timesteps sinusoid.

The reference has no tests; diffusers==0.24.0 is not installable here.  These constants were
derived from the published algorithm (rows a2/a4 of SURVEY.md §8) and pin both the oracle
restatement and the product's host-side scheduler; they also detect the wrong-schedule bug
(training's ``scaled_linear`` gives alpha_bar[981]=0.005775...).
"""
import numpy as np
import pytest
import torch

from oracle.diffusers_restated import DDIMSchedulerRef, get_timestep_embedding
from oracle.loop_ref import make_scheduler

ABAR = {0: 0.999149978, 1: 0.998289526, 21: 0.978936434, 501: 0.159735978, 901: 0.004885187,
        951: 0.002782951, 981: 0.001958828, 999: 0.001578963}


def schedulers():
    from rcdms_b200.schedulers import DDIMScheduler
    from rcdms_b200.unet_spec import RCDMS_SCHEDULER_KWARGS
    prod = DDIMScheduler(**RCDMS_SCHEDULER_KWARGS, steps_offset=1, clip_sample=False)
    return [("oracle", make_scheduler()), ("product", prod)]


@pytest.mark.parametrize("which", [0, 1])
def test_alphas_cumprod(which):
    name, s = schedulers()[which]
    for i, v in ABAR.items():
        assert abs(float(s.alphas_cumprod[i]) - v) < 2e-9 + 1e-6 * v, (name, i)
    assert float(s.alphas_cumprod[981]).hex().startswith("0x1.00bf5e"), name


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("n,head,tail", [(50, [981, 961, 941, 921], [41, 21, 1]), (20, [951, 901], [51, 1]),
                                         (10, [901, 801], [101, 1])])
def test_timesteps_bit_identical(which, n, head, tail):
    name, s = schedulers()[which]
    s.set_timesteps(n)
    ts = s.timesteps
    assert ts.dtype == torch.int64 and len(ts) == n, name
    assert ts[: len(head)].tolist() == head and ts[-len(tail):].tolist() == tail, name
    assert ts.tolist() == ((np.arange(n) * (1000 // n))[::-1] + 1).tolist()


@pytest.mark.parametrize("which", [0, 1])
def test_one_step(which):
    name, s = schedulers()[which]
    s.set_timesteps(50)
    x = torch.ones(4)
    eps = torch.full((4,), 0.5)
    out = s.step(eps, 981, x, eta=0.0)
    assert torch.allclose(out.prev_sample, torch.full((4,), 1.0623393), atol=2e-6), name
    assert torch.allclose(out.pred_original_sample, torch.full((4,), 11.3082972), atol=2e-5), name
    last = s.step(eps, 1, x, eta=0.0).prev_sample  # prev_timestep < 0 -> final_alpha_cumprod = 1
    assert torch.allclose(last, torch.full((4,), 0.9801597), atol=2e-6), name
    assert s.init_noise_sigma == 1.0 and s.order == 1
    assert s.scale_model_input(x, 981) is x


def test_wrong_schedule_is_detectable():
    s = DDIMSchedulerRef(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    assert abs(float(s.alphas_cumprod[981]) - 0.005775496) < 1e-7


def test_timestep_sinusoid():
    e = get_timestep_embedding(torch.tensor([981]), 320, flip_sin_to_cos=True, downscale_freq_shift=0)[0]
    assert torch.allclose(e[0:3], torch.tensor([0.67995721, -0.79842919, 0.57806414]), atol=2e-4)
    assert torch.allclose(e[160:163], torch.tensor([0.73325181, 0.60208869, 0.81599128]), atol=2e-4)
