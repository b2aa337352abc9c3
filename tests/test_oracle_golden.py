"""Pin the oracle restatement (oracle/unet_ref.py) to the REFERENCE modules.

tests/golden/unet_*.pt were produced by the reference's own src/models/unet.py (imported
unmodified through oracle/diffusers_shim) by ``python -m oracle.make_golden``.
"""
import json
import os

import pytest
import torch

from oracle.unet_ref import unet_forward
from rcdms_b200.synthetic import synthetic_state_dict
from rcdms_b200.unet_spec import full_config, state_dict_spec, tiny_config

from conftest import GOLDEN

CASES = [("tiny_8x8", tiny_config), ("tiny_16x16", tiny_config), ("full_8x8", full_config)]


def golden_inputs(cfg, b, f, h, w, L, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((b, cfg["in_channels"], f, h, w), generator=g)
    ctx = torch.randn((b * f, L, cfg["cross_attention_dim"]), generator=g)
    return x, ctx


@pytest.mark.parametrize("name,cfg_fn", CASES)
def test_oracle_matches_reference_golden(name, cfg_fn):
    cfg = cfg_fn()
    gold = torch.load(os.path.join(GOLDEN, f"unet_{name}.pt"))
    b, f, h, w, L = gold["shape"]
    sd = synthetic_state_dict(cfg, seed=gold["weight_seed"])
    x, ctx = golden_inputs(cfg, b, f, h, w, L, gold["input_seed"])
    taps = {}
    with torch.no_grad():
        y = unet_forward(sd, cfg, x, gold["timestep"], ctx, taps=taps)
    assert y.shape == gold["out"].shape
    err = (y - gold["out"]).abs().max().item()
    assert err < 2e-5, err  # fp32 CPU, same op order: only BLAS blocking noise
    for k, v in gold["taps"].items():
        mine = taps[k][:1, :8, :, :4, :4]
        assert (mine - v).abs().max().item() < 2e-4 * max(1.0, v.abs().max().item()), k


def test_state_dict_spec_matches_reference():
    spec = json.load(open(os.path.join(GOLDEN, "state_dict_spec.json")))
    mine = state_dict_spec(full_config())
    assert len(mine) == 1286 == len(spec)
    assert [[n, list(s)] for n, s in mine] == spec
    n_params = sum(int(torch.Size(s).numel()) for _, s in mine)
    assert abs(n_params / 1e6 - 1276.88) < 0.01
