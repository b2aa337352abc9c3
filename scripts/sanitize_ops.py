"""Small-shape run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): tcgen05 GEMM (1-CTA,
CTA pair, stream-K on, LayerNorm fold, GEGLU, GroupNorm statistics epilogue), implicit-GEMM convs (stride 1 / 2 / asymmetric /
folded upsample), GroupNorm (grid-barrier kernel, two-kernel path, statistics-driven apply), LayerNorm, flash attention,
temporal attention, CFG + DDIM step, the small VAE kernels, and one tiny UNet forward + 2-step CUDA-graph denoise loop.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_ops.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import parity_checks as pc  # noqa: E402
from rcdms_b200 import _lib  # noqa: E402

L = _lib.lib()
f16 = torch.float16
res = []


def run(name, fn):
    r = fn()
    torch.cuda.synchronize()
    ok = r["ok"] if isinstance(r, dict) else bool(r)
    res.append((name, ok))
    print(("ok     " if ok else "FAILED ") + name, flush=True)


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
if ONLY == "ffn":  # (re-run of a single family: python scripts/sanitize_ops.py ffn)
    run("ffn fused 256 rows f16", lambda: pc.check_ffn_fused(256, f16))
    run("ffn fused 200 rows bf16", lambda: pc.check_ffn_fused(200, torch.bfloat16))
    print("ALL OK" if all(ok for _, ok in res) else "SOME FAILED", len(res), "checks")
    sys.exit(0)
if ONLY == "xattn":  # the short-key-range attention kernel alone: python scripts/sanitize_ops.py xattn
    for (b, hds, sq, skv, d) in [(2, 8, 256, 85, 40), (2, 8, 64, 91, 80), (2, 8, 64, 64, 160), (2, 8, 100, 112, 40),
                                 (3, 4, 33, 5, 64), (2, 8, 16, 7, 16)]:
        run(f"short-kv mma.sync attention b{b} h{hds} Sq{sq} Skv{skv} d{d}", lambda: pc.check_flash(b, hds, sq, skv, d, f16, short_kv=1))
    print("ALL OK" if all(ok for _, ok in res) else "SOME FAILED", len(res), "checks")
    sys.exit(0)
if ONLY == "folds":  # the last session's kernels / launch forms: python scripts/sanitize_ops.py folds
    import test_prior_gpu as tp  # noqa: E402  (assert-based checks, called directly)

    def asserts(fn, *a):
        fn(*a)
        return True
    run("LN-folded chain 170x128 -> 384, 5 frames + pe", lambda: asserts(tp.test_layernorm_folded_gemm_chain, 170, 128, 384, None, False, 5, 17, f16))
    run("LN-folded chain 170x128 -> 512 gelu", lambda: asserts(tp.test_layernorm_folded_gemm_chain, 170, 128, 512, "gelu", False, 1, 17, f16))
    run("LN-folded chain 170x128 -> 1024 geglu", lambda: asserts(tp.test_layernorm_folded_gemm_chain, 170, 128, 1024, "geglu", False, 1, 17, f16))
    run("LN-folded chain 640x1280 -> 1280 + res (192-wide producer parts)", lambda: asserts(tp.test_layernorm_folded_gemm_chain, 640, 1280, 1280, None, True, 1, 64, f16))
    run("proj_out over ff2 170x128", lambda: asserts(tp.test_proj_out_folded_over_ff2, 170, 128, f16))
    run("proj_out over ff2 640x320 bf16", lambda: asserts(tp.test_proj_out_folded_over_ff2, 640, 320, torch.bfloat16))
    for (b, hds, sq, skv, d) in [(2, 8, 256, 85, 40), (2, 8, 100, 112, 40), (2, 8, 64, 91, 80), (2, 8, 64, 64, 160), (3, 4, 33, 5, 64),
                                 (2, 8, 16, 7, 16)]:
        run(f"short-kv mma.sync attention b{b} h{hds} Sq{sq} Skv{skv} d{d}", lambda: pc.check_flash(b, hds, sq, skv, d, f16, short_kv=1))
    import unet_checks as uc  # noqa: E402
    from rcdms_b200.unet_spec import tiny_config  # noqa: E402
    r = uc.run_case(tiny_config(), (2, 5, 8, 8, 7), 981, f16)
    run("tiny unet forward (two-segment proj_out launches)", lambda: bool(
        r["stats"]["finite"] and r["stats"]["max_abs"] <= max(3 * r["floor"]["max_abs"], 5e-3)))
    from rcdms_b200.prior_spec import prior_tiny_config  # noqa: E402
    pr = tp._forward_case(prior_tiny_config(), f16, 500)
    run("tiny prior forward (folded LayerNorms, rcdm_gemm_cat)", lambda: bool(pr["max"] <= max(3 * pr["fmax"], 5e-3)))
    print("ALL OK" if all(ok for _, ok in res) else "SOME FAILED", len(res), "checks")
    sys.exit(0)
for pair, sk in ((0, 0), (2, 1)):
    L.rcdm_set_gemm_pair(pair)
    L.rcdm_set_stream_k_min(sk)
    tag = f"[pair{pair} sk{sk}] "
    run(tag + "linear 640x320x320 +res", lambda: pc.check_linear(640, 320, 320, f16, residual=True))
    run(tag + "linear 2560x1280x1280", lambda: pc.check_linear(2560, 1280, 1280, f16, residual=True))
    run(tag + "geglu 640x640", lambda: pc.check_geglu(640, 640, f16))
    run(tag + "linear_ln 640x960x320 pe", lambda: pc.check_linear_ln(640, 960, 320, f16, pe=True))
    run(tag + "rowstats 640x320x320", lambda: pc.check_rowstats(640, 320, 320, f16))
    run(tag + "conv3x3 s1", lambda: pc.check_conv3x3(10, 8, 8, 320, 320, 1, f16, residual=True))
    run(tag + "conv3x3 s2", lambda: pc.check_conv3x3(5, 16, 16, 192, 320, 2, f16))
    run(tag + "upsample_conv3x3", lambda: pc.check_upsample_conv3x3(10, 8, 8, 320, 320, f16))
    run(tag + "groupnorm_from_stats", lambda: pc.check_groupnorm_from_stats(10, 64, 5, 640, 320, 64, f16, residual=True))
L.rcdm_set_gemm_pair(1)
L.rcdm_set_stream_k_min(24)
run("groupnorm fused (grid barrier)", lambda: pc.check_groupnorm(2, 5 * 64, 320, f16))
run("groupnorm per frame", lambda: pc.check_groupnorm(10, 64, 640, f16, silu=False, eps=1e-6))
L.rcdm_debug_set_option(b"gn_fused", 0)
run("groupnorm two-kernel path", lambda: pc.check_groupnorm(2, 5 * 64, 320, f16))
L.rcdm_debug_set_option(b"gn_fused", 1)
run("layernorm + pe", lambda: pc.check_layernorm(640, 320, f16, pe=True))
run("flash self d40", lambda: pc.check_flash(2, 8, 256, 256, 40, f16))
run("flash cross L85 d80", lambda: pc.check_flash(2, 8, 256, 85, 80, f16))
run("flash d160", lambda: pc.check_flash(2, 8, 64, 64, 160, f16))
run("temporal d40", lambda: pc.check_temporal(2, 5, 64, 8, 40, f16))
run("ddim cfg step", lambda: pc.check_ddim(1, 5, 8, 8, f16))
run("context fusion", lambda: pc.check_context_fusion(2, 7, 96, 9, 16, 96, 8, f16))
run("ffn fused 256 rows", lambda: pc.check_ffn_fused(256, f16))

# VAE pieces + tiny decode / encode
from rcdms_b200.models import AutoencoderKL  # noqa: E402
from rcdms_b200.vae_spec import synthetic_vae_state_dict, vae_tiny_config  # noqa: E402
vc = vae_tiny_config()
vae = AutoencoderKL.from_config(vc)
vae.load_state_dict(synthetic_vae_state_dict(vc))
vae = vae.to("cuda", f16)
img = vae.decode(torch.randn((2, 4, 8, 8), device="cuda", dtype=f16)).sample
mom = vae.encode(torch.randn((1, 3, 32, 32), device="cuda", dtype=f16)).latent_dist.parameters
run("vae tiny decode + encode", lambda: bool(torch.isfinite(img).all() and torch.isfinite(mom).all()))

# tiny UNet forward + CUDA-graph loop
import unet_checks as uc  # noqa: E402
from rcdms_b200.unet_spec import tiny_config  # noqa: E402
r = uc.run_case(tiny_config(), (2, 5, 8, 8, 7), 981, f16)
run("tiny unet forward", lambda: bool(r["stats"]["finite"] and r["stats"]["max_abs"] <= max(3 * r["floor"]["max_abs"], 5e-3)))
print("ALL OK" if all(ok for _, ok in res) else "SOME FAILED", len(res), "checks")
