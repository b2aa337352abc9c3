"""A few GEMM / conv launches for `ncu --set full` (3 launches per configuration; capture the last of each).
Order: conv 64x64 640->320 [1cta], same [pair], gemm 10240x640x640 [1cta], same [pair], conv 16x16 1280->1280 [pair+sk]"""
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


def gemm(M, N, K, reps=3):
    a = torch.randn((M, K), device="cuda").to(dt)
    w = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(dt)
    b = torch.randn((N,), device="cuda")
    r = torch.randn((M, N), device="cuda").to(dt)
    out = torch.empty((M, N), dtype=dt, device="cuda")
    for _ in range(reps):
        _lib.check(L.rcdm_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr(), out.data_ptr(), M, N, K, 0, 0, 0, s))
    torch.cuda.synchronize()


def conv(n, h, cin, cout, reps=3):
    x = torch.randn((n, h, h, cin), device="cuda").to(dt)
    wp = (torch.randn((cout, 9 * cin), device="cuda") / math.sqrt(9 * cin)).to(dt)
    b = torch.randn((cout,), device="cuda")
    out = torch.empty((n, h, h, cout), dtype=dt, device="cuda")
    for _ in range(reps):
        _lib.check(L.rcdm_conv3x3(1, x.data_ptr(), wp.data_ptr(), b.data_ptr(), None, out.data_ptr(), n, h, h, cin, cout, 1, 0, s))
    torch.cuda.synchronize()


L.rcdm_set_stream_k_min(0)
for pair in (0, 2):
    L.rcdm_set_gemm_pair(pair)
    conv(10, 64, 640, 320)
for pair in (0, 2):
    L.rcdm_set_gemm_pair(pair)
    gemm(10240, 640, 640)
L.rcdm_set_stream_k_min(24)
conv(10, 16, 1280, 1280)
