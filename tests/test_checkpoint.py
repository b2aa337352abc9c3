"""Checkpoint ingestion: the DeepSpeed ``["module"]`` prefix split of
stage2_batchtest_rcdms_model.py:225-243 / stage1_batchtest_rcdms_model.py:102-103 into the drop-in modules, and the
offline flat weight file."""
import pytest
import torch

from rcdms_b200 import checkpoint as ck
from rcdms_b200.models import MyPriorTransformer, UNet3DConditionModel
from rcdms_b200.pipelines.RCDMs_pipeline import local_feature
from rcdms_b200.prior_spec import prior_tiny_config
from rcdms_b200.synthetic import synthetic_prior_state_dict, synthetic_state_dict
from rcdms_b200.unet_spec import tiny_config


def _deepspeed_file(tmp_path, cfg):
    torch.manual_seed(0)
    lm = local_feature(text_dim=96, vis_dim=16, hidden_dim=96, num_heads=8)
    gm = local_feature(text_dim=96, vis_dim=12, hidden_dim=96, num_heads=8)
    unet_sd = synthetic_state_dict(cfg, seed=3)
    module = {}
    module.update({"seen_module." + k: v.clone() for k, v in lm.state_dict().items()})
    module.update({"unseen_module." + k: v.clone() for k, v in gm.state_dict().items()})
    module.update({"unet." + k: v for k, v in unet_sd.items()})
    module["stray.key"] = torch.zeros(1)
    path = tmp_path / "mp_rank_00_model_states.pt"
    torch.save({"module": module, "epoch": 3, "last_global_step": 1234}, path)  # train_stage2.py:60-77 client state
    return str(path), lm, gm, unet_sd


def test_stage2_split_and_strict_load(tmp_path):
    cfg = tiny_config()
    path, lm, gm, unet_sd = _deepspeed_file(tmp_path, cfg)
    unet = UNet3DConditionModel.from_config(cfg)
    lm2 = local_feature(text_dim=96, vis_dim=16, hidden_dim=96, num_heads=8)
    gm2 = local_feature(text_dim=96, vis_dim=12, hidden_dim=96, num_heads=8)
    other = ck.load_stage2_checkpoint(path, unet, lm2, gm2)
    assert other == ["stray.key"]
    assert all(torch.equal(v, unet_sd[k]) for k, v in unet.state_dict().items())
    assert all(torch.equal(v, lm.state_dict()[k]) for k, v in lm2.state_dict().items())
    assert all(torch.equal(v, gm.state_dict()[k]) for k, v in gm2.state_dict().items())
    # strictness: a missing UNet entry is an error, as in the reference's load_state_dict
    sd = torch.load(path)["module"]
    sd.pop("unet.conv_in.weight")
    with pytest.raises(RuntimeError):
        ck.load_stage2_checkpoint(sd, UNet3DConditionModel.from_config(cfg))
    seen, unseen, un, oth = ck.split_stage2_module_state({"unet.a.unet.b": 1, "seen_module.x": 2, "unseen_module.y": 3,
                                                          "unetx": 4, "zzz": 5})
    assert un == {"a.b": 1, "unetx": 4} and seen == {"x": 2} and unseen == {"y": 3} and oth == ["zzz"]  # str.replace quirk


def test_stage1_load(tmp_path):
    cfg = prior_tiny_config()
    sd = synthetic_prior_state_dict(cfg, seed=5)
    path = tmp_path / "prior_states.pt"
    torch.save({"module": sd}, path)
    m = MyPriorTransformer.from_config(cfg)
    ck.load_stage1_checkpoint(str(path), m)
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, None])
def test_flat_file_round_trip(tmp_path, dtype):
    cfg = tiny_config()
    path, lm, gm, unet_sd = _deepspeed_file(tmp_path, cfg)
    flat = str(tmp_path / "stage2.rcdmflat")
    size = ck.convert_stage2_checkpoint(path, flat, dtype) if dtype is not None else ck.save_flat(
        torch.load(path)["module"], flat, None)
    module = torch.load(path)["module"]
    back = ck.load_flat(flat)
    assert list(back) == list(module)
    n_el = sum(v.numel() for v in module.values())
    assert size < n_el * (2 if dtype is not None else 4) + 256 * (len(module) + 8) + 200 * len(module)
    for k, v in module.items():
        want = v.to(dtype) if (dtype is not None and v.is_floating_point()) else v
        assert back[k].dtype == want.dtype and back[k].shape == want.shape
        assert torch.equal(back[k], want), k
    # the flat file feeds the same loader (keys keep their prefixes)
    unet = UNet3DConditionModel.from_config(cfg)
    assert ck.load_stage2_checkpoint(flat, unet) == ["stray.key"]
    ref = {k: (v.to(dtype).float() if dtype is not None else v) for k, v in unet_sd.items()}
    assert all(torch.equal(v, ref[k]) for k, v in unet.state_dict().items())
    with pytest.raises(ValueError):
        ck.load_flat(path)


@pytest.mark.gpu
def test_checkpoint_to_forward_on_gpu(tmp_path):
    """A split DeepSpeed ``["module"]`` file (and its offline flat-file conversion) loaded through
    ``load_stage2_checkpoint`` must reproduce the forward of the directly loaded model bit for bit on the device
    (``stage2_batchtest_rcdms_model.py:225-243`` then ``RCDMs_pipeline.py:488``)."""
    cfg = tiny_config()
    path, lm, gm, unet_sd = _deepspeed_file(tmp_path, cfg)
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, cfg["in_channels"], 5, 16, 16), generator=g).cuda().half()
    ctx = torch.randn((10, 7, cfg["cross_attention_dim"]), generator=g).cuda().half()

    def forward(m):
        m = m.to(device="cuda", dtype=torch.float16)
        return m(x, 501, encoder_hidden_states=ctx, return_dict=False)[0]

    direct = UNet3DConditionModel.from_config(cfg)
    direct.load_state_dict(unet_sd, strict=True)
    y_direct = forward(direct)
    assert torch.isfinite(y_direct).all() and y_direct.abs().mean().item() > 1e-3

    from_ds = UNet3DConditionModel.from_config(cfg)
    lm2 = local_feature(text_dim=96, vis_dim=16, hidden_dim=96, num_heads=8)
    gm2 = local_feature(text_dim=96, vis_dim=12, hidden_dim=96, num_heads=8)
    assert ck.load_stage2_checkpoint(path, from_ds, lm2, gm2) == ["stray.key"]
    assert torch.equal(forward(from_ds), y_direct)

    flat = str(tmp_path / "stage2.rcdmflat")
    ck.convert_stage2_checkpoint(path, flat, torch.float16)
    from_flat = UNet3DConditionModel.from_config(cfg)
    ck.load_stage2_checkpoint(flat, from_flat, local_feature(96, 16, 96, 8), local_feature(96, 12, 96, 8))
    assert torch.equal(forward(from_flat), y_direct)

    # a checkpoint with different weights must give a different output (the loader is not a no-op)
    other = UNet3DConditionModel.from_config(cfg)
    other.load_state_dict(synthetic_state_dict(cfg, seed=4), strict=True)
    assert not torch.equal(forward(other), y_direct)
