"""Quick correctness smoke of the CTA-pair GEMM (run first, under a short timeout, before the full GPU round)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import parity_checks as pc  # noqa: E402
from rcdms_b200 import _lib  # noqa: E402

L = _lib.lib()
for pair in (0, 2):
    L.rcdm_set_gemm_pair(pair)
    for sk in (0, 1):
        L.rcdm_set_stream_k_min(sk)
        for thunk in (lambda: pc.check_linear(512, 320, 256, torch.float16),
                      lambda: pc.check_linear(640, 320, 320, torch.float16, residual=True),
                      lambda: pc.check_linear(2560, 1280, 1280, torch.float16, residual=True),
                      lambda: pc.check_linear(1000, 128, 96, torch.bfloat16),
                      lambda: pc.check_geglu(640, 640, torch.float16),
                      lambda: pc.check_conv3x3(10, 16, 16, 320, 640, 1, torch.float16),
                      lambda: pc.check_conv3x3(10, 8, 8, 1280, 640, 1, torch.float16),
                      lambda: pc.check_conv3x3(5, 16, 16, 192, 320, 2, torch.float16)):
            r = thunk()
            torch.cuda.synchronize()
            print("pair", pair, "sk", sk, "PASS" if r["ok"] else "FAIL", r["name"], "%.2e" % r["max_abs"], flush=True)
