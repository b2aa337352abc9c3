"""Parity + timing of the fused GEGLU feed-forward kernel against the two-kernel path (diagnostic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import parity_checks as pc  # noqa: E402
from rcdms_b200 import _lib, ops  # noqa: E402

if os.environ.get("FFN_MODE"):  # 1 = single-CTA kernel, 2 = CTA-pair kernel
    _lib.lib().rcdm_debug_set_option(b"ffn_fused", int(os.environ["FFN_MODE"]))

for dt in (torch.float16, torch.bfloat16):
    for M in (128, 1000, 20480):
        r = pc.check_ffn_fused(M, dt)
        print({k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
if len(sys.argv) > 1:
    import math
    M, C, J = 40960, 320, 1280
    dt = torch.float16
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.randn((M, C), device="cuda", generator=g).to(dt)
    w1 = (torch.randn((2 * J, C), device="cuda", generator=g) / math.sqrt(C)).to(dt)
    b1 = torch.randn((2 * J,), device="cuda", generator=g)
    gamma = torch.ones((C,), device="cuda")
    beta = torch.zeros((C,), device="cuda")
    w2 = (torch.randn((C, J), device="cuda", generator=g) / math.sqrt(J)).to(dt)
    b2 = torch.randn((C,), device="cuda", generator=g)

    def timeit(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    print("fused (incl. per-call fold / pack / rowstats launches): %.1f us" % timeit(lambda: ops.ffn_geglu_ln(y, w1, b1, gamma, beta, w2, b2)))
    mid = ops.linear_ln(y, w1, gamma, beta, b1, None, 1, True)
    print("two kernels, GEGLU GEMM (incl. the same per-call launches): %.1f us" % timeit(lambda: ops.linear_ln(y, w1, gamma, beta, b1, None, 1, True)))
    print("two kernels, second GEMM: %.1f us" % timeit(lambda: ops.linear(mid, w2, b2, y)))
