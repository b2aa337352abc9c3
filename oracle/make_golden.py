"""TEST INFRASTRUCTURE ONLY (oracle) — generate tests/golden/*.pt from the REFERENCE modules.

Run in the build container (needs /root/reference; cannot run on the GPU box):

    python -m oracle.make_golden

Imports the reference's own ``src/models/unet.py`` UNCHANGED through ``oracle/diffusers_shim``,
loads the deterministic synthetic state dict (``rcdms_b200.synthetic``), runs fp32 CPU forwards
and saves small input/output/intermediate tensors.  The committed fixtures pin
``oracle/unet_ref.py`` (and through it the CUDA path) to the reference implementation.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root() -> str:
    """Where the unmodified reference lives: RCDMS_REFERENCE, else a driver-installed baseline/_ref, else /root/reference."""
    for cand in (os.environ.get("RCDMS_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "src", "models", "unet.py")):
            return cand
    raise RuntimeError("reference checkout not found (set RCDMS_REFERENCE)")


def load_reference_module(name: str):
    """Import ``<reference>/src/models/<name>.py`` UNCHANGED under the private package ``_rcdms_reference.models``.

    The repo ships its own ``src/`` import shim (the drop-in boundary), so a plain ``from src.models.unet import ...``
    can resolve to the product instead of the reference depending on ``sys.path`` order and on what is cached in
    ``sys.modules``.  A private package name cannot be shadowed; the reference's files use relative imports only
    (``from .unet_blocks import ...``), so they load unchanged.  The result is asserted to come from the reference tree."""
    import importlib
    import inspect
    import types
    ref = reference_root()
    shim = os.path.join(ROOT, "oracle", "diffusers_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    for pkg, sub in (("_rcdms_reference", "src"), ("_rcdms_reference.models", os.path.join("src", "models"))):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(ref, sub)]
            sys.modules[pkg] = m
    mod = importlib.import_module(f"_rcdms_reference.models.{name}")
    src = os.path.realpath(inspect.getsourcefile(mod))
    if not src.startswith(os.path.realpath(ref) + os.sep):
        raise RuntimeError(f"golden recipe imported {src}, not the reference under {ref}")
    return mod


def load_reference_unet(cfg):
    UNet3DConditionModel = load_reference_module("unet").UNet3DConditionModel  # the reference's class, unmodified
    init = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    return UNet3DConditionModel.from_config(init)


def golden_inputs(cfg, b, f, h, w, L, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((b, cfg["in_channels"], f, h, w), generator=g)
    ctx = torch.randn((b * f, L, cfg["cross_attention_dim"]), generator=g)
    return x, ctx


def load_reference_prior(cfg):
    MyPriorTransformer = load_reference_module("myprior_transformer").MyPriorTransformer  # unmodified reference class
    init = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    return MyPriorTransformer.from_config(init)


def prior_goldens(out_dir):
    """Stage-1 prior: reference MyPriorTransformer forwards on the synthetic state dict."""
    from rcdms_b200.prior_spec import prior_tiny_config
    from rcdms_b200.synthetic import synthetic_prior_inputs, synthetic_prior_state_dict
    cases = [
        ("prior_tiny", prior_tiny_config(), 500),
        ("prior_tiny_norms", prior_tiny_config(norm_in_type="layer", embedding_proj_norm_type="layer", num_layers=1,
                                               added_emb_type=None, additional_embeddings=5), 17),
        ("prior_wide", prior_tiny_config(num_attention_heads=8, num_layers=1, embedding_dim=96, num_embeddings=27), 999),
    ]
    for name, cfg, t in cases:
        model = load_reference_prior(cfg).eval()
        sd = synthetic_prior_state_dict(cfg, seed=0)
        model.load_state_dict(sd, strict=True)
        inp = synthetic_prior_inputs(cfg, clip_index=3)
        x = torch.cat([inp["latents"]] * 2)
        taps, hooks = {}, []
        for mod_name in ("transformer_blocks.0", "transformer_blocks.1"):
            mod = dict(model.named_modules())[mod_name]
            hooks.append(mod.register_forward_hook(lambda m, i, o, n=mod_name: taps.__setitem__(n, o.detach().clone())))
        with torch.no_grad():
            out = model(x, torch.tensor(t), inp["prompt_embeds"], inp["text_hidden"],
                        torch.cat([inp["imgs_proj_embeds1"]] * 2), torch.cat([inp["mask_label"]] * 2),
                        inp["text_mask"]).predicted_image_embedding
            for hk in hooks:
                hk.remove()
            out_nomask = model(x, t, inp["prompt_embeds"], inp["text_hidden"], torch.cat([inp["imgs_proj_embeds1"]] * 2),
                               torch.cat([inp["mask_label"]] * 2), None, return_dict=False)[0]
        torch.save(dict(cfg=cfg, timestep=t, out=out.clone(), out_nomask=out_nomask.clone(),
                        taps={k: v[:, -2:].clone() for k, v in taps.items()},
                        state_dict_names=[(k, tuple(v.shape)) for k, v in model.state_dict().items()]),
                   os.path.join(out_dir, f"{name}.pt"))
        print(name, tuple(out.shape), float(out.abs().mean()))


def main():
    from rcdms_b200.synthetic import synthetic_state_dict
    from rcdms_b200.unet_spec import full_config, tiny_config
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if "prior" in sys.argv[1:] or not sys.argv[1:]:
        prior_goldens(out_dir)
        if "prior" in sys.argv[1:]:
            return
    torch.manual_seed(0)
    cases = [
        # name, cfg, (b, f, h, w, L), timestep, taps kept
        ("tiny_8x8", tiny_config(), (2, 5, 8, 8, 7), 981),
        ("tiny_16x16", tiny_config(), (2, 5, 16, 16, 85), 501),
        ("full_8x8", full_config(), (2, 5, 8, 8, 85), 981),
    ]
    for name, cfg, (b, f, h, w, L), t in cases:
        model = load_reference_unet(cfg).eval()
        sd = synthetic_state_dict(cfg, seed=0)
        model.load_state_dict(sd, strict=True)
        x, ctx = golden_inputs(cfg, b, f, h, w, L, seed=1234)
        taps = {}
        hooks = []
        for mod_name in ("conv_in", "down_blocks.0.resnets.0", "down_blocks.0.attentions.0",
                         "down_blocks.0.motion_modules.0", "down_blocks.0.downsamplers.0", "mid_block",
                         "up_blocks.0.upsamplers.0", "up_blocks.3.motion_modules.2"):
            mod = model.get_submodule(mod_name)

            def hook(m, i, o, key=mod_name):
                o = o[0] if isinstance(o, tuple) else (o.sample if hasattr(o, "sample") else o)
                taps[key] = o.detach().clone()
            hooks.append(mod.register_forward_hook(hook))
        with torch.no_grad():
            y = model(x, torch.tensor(t), encoder_hidden_states=ctx, return_dict=False)[0]
        for hk in hooks:
            hk.remove()
        keep = {k: v[:1, :8, :, :4, :4].contiguous() for k, v in taps.items()}  # small corner of each tap
        torch.save(dict(shape=(b, f, h, w, L), timestep=t, input_seed=1234, weight_seed=0,
                        out=y.contiguous(), taps=keep), os.path.join(out_dir, f"unet_{name}.pt"))
        print(name, tuple(y.shape), float(y.abs().mean()), float(y.abs().max()))
        del model, sd


if __name__ == "__main__":
    main()
