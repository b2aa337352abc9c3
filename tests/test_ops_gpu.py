"""GPU parity of every CUDA kernel, called through the C ABI (ctypes), vs an fp32 torch reference fed the same
16-bit-rounded inputs.  Tolerances (written here): fp16 rtol 1e-3 / atol 1e-3*rms(ref) per op
(x2 for norm/GEGLU epilogues, x3 for attention, which chain two roundings); bf16 rtol 1.6e-2."""
import pytest
import torch

import parity_checks as pc

pytestmark = pytest.mark.gpu

_CHECKS = list(pc.all_op_checks()) if torch.cuda.is_available() else []


@pytest.mark.parametrize("idx", range(len(_CHECKS)))
def test_op_parity(idx):
    r = _CHECKS[idx]()
    assert r["ok"], r


def test_native_library_is_loaded_and_counts_launches():
    from rcdms_b200 import _lib
    L = _lib.lib()
    before = L.rcdm_kernel_launches()
    pc.check_linear(256, 160, 64, torch.float16)
    assert L.rcdm_kernel_launches() > before
    assert L.rcdm_device_count() >= 1


def test_errors_are_python_exceptions():
    from rcdms_b200 import _lib, ops
    a = torch.zeros((8, 12), dtype=torch.float16, device="cuda")   # K = 12 is not a multiple of 8
    w = torch.zeros((16, 12), dtype=torch.float16, device="cuda")
    with pytest.raises(_lib.RcdmError):
        ops.linear(a, w)
    with pytest.raises(TypeError):
        ops.linear(a.float(), w.float())


@pytest.mark.parametrize("pair,sk", [(0, 0), (2, 0), (0, 1), (2, 1)])
def test_gemm_scheduling_variants_agree(pair, sk):
    """CTA-pair (cta_group::2) kernel and stream-K decomposition forced on / off through the tuning knobs: every
    combination must pass the same parity checks, and without stream-K pairing must not change a single bit (same K
    order per output element; with stream-K the split points depend on the worker count, so only parity is required)."""
    from rcdms_b200 import _lib, ops
    L = _lib.lib()
    prev_pair, prev_sk = L.rcdm_set_gemm_pair(pair), L.rcdm_set_stream_k_min(sk)
    try:
        for thunk in (lambda: pc.check_linear(640, 320, 320, torch.float16, residual=True),
                      lambda: pc.check_linear(2560, 1280, 1280, torch.float16, residual=True),
                      lambda: pc.check_linear(1000, 128, 96, torch.bfloat16),
                      lambda: pc.check_geglu(640, 640, torch.float16),
                      lambda: pc.check_linear_ln(640, 960, 320, torch.float16, pe=True),
                      lambda: pc.check_rowstats(2560, 1280, 1280, torch.float16),
                      lambda: pc.check_conv3x3(10, 16, 16, 320, 640, 1, torch.float16),
                      lambda: pc.check_conv3x3(10, 8, 8, 1280, 640, 1, torch.float16),
                      lambda: pc.check_conv3x3(5, 16, 16, 192, 320, 2, torch.float16)):
            r = thunk()
            assert r["ok"], (pair, sk, r)
        g = torch.Generator(device="cuda").manual_seed(11)
        a = torch.randn((2560, 640), generator=g, device="cuda").half()
        w = (torch.randn((1280, 640), generator=g, device="cuda") / 25).half()
        out = ops.linear(a, w)
        L.rcdm_set_gemm_pair(0)
        ref = ops.linear(a, w)
        if sk == 0:
            assert torch.equal(out, ref), "pairing changed the result bitwise"
        else:
            assert (out.float() - ref.float()).abs().max().item() <= 2e-3 * ref.float().abs().max().item()
    finally:
        L.rcdm_set_gemm_pair(prev_pair)
        L.rcdm_set_stream_k_min(prev_sk)


@pytest.mark.parametrize("mode", [1, 2])
def test_fused_feed_forward_kernels(mode):
    """Both builds of the opt-in fused GEGLU feed-forward (library option "ffn_fused": 1 = single-CTA kernel, 2 = CTA-pair
    kernel; the standalone entry point takes the kernel from the option) against fp32 and the two-GEMM path."""
    from rcdms_b200 import _lib
    L = _lib.lib()
    prev = L.rcdm_debug_set_option(b"ffn_fused", mode)
    try:
        for M, dt in ((128, torch.float16), (1000, torch.float16), (4224, torch.bfloat16)):
            r = pc.check_ffn_fused(M, dt)
            assert r["ok"], (mode, r)
    finally:
        L.rcdm_debug_set_option(b"ffn_fused", prev)
