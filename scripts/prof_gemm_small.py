"""K = 320 GEMM launches for `ncu --set full --import-source on` (3 launches per shape; capture the last of each):
(40960, 320, 320) + residual, (40960, 960, 320), GEGLU (40960, 2560, 320), and the two-segment launch that ends every
320-channel transformer / motion module since proj_out is folded over ff.net.2: (40960, 320, 320 + 1280) + residual."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200 import _lib  # noqa: E402

dt = torch.float16
L = _lib.lib()
s = _lib.current_stream_ptr()


def gemm(M, N, K, res, geglu=0, reps=3):
    a = torch.randn((M, K), device="cuda").to(dt)
    w = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(dt)
    b = torch.randn((N,), device="cuda")
    r = torch.randn((M, N), device="cuda").to(dt) if res else None
    out = torch.empty((M, N // 2 if geglu else N), dtype=dt, device="cuda")
    for _ in range(reps):
        _lib.check(L.rcdm_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr() if res else None, out.data_ptr(),
                               M, N, K, geglu, 0, 0, s))
    torch.cuda.synchronize()


gemm(40960, 320, 320, True)
gemm(40960, 960, 320, False)
gemm(40960, 2560, 320, False, 1)


def gemm_cat(M, C, reps=3):
    y = torch.randn((M, C), device="cuda").to(dt)
    g = torch.randn((M, 4 * C), device="cuda").to(dt)
    w = (torch.randn((C, 5 * C), device="cuda") / math.sqrt(5 * C)).to(dt)
    b = torch.randn((C,), device="cuda")
    x = torch.randn((M, C), device="cuda").to(dt)
    for _ in range(reps):
        _lib.check(L.rcdm_gemm_cat(1, y.data_ptr(), C, g.data_ptr(), 4 * C, w.data_ptr(), b.data_ptr(), x.data_ptr(),
                                   x.data_ptr(), M, C, None, s))
    torch.cuda.synchronize()


gemm_cat(40960, 320)
