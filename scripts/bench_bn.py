"""Tile-width sweep of the plain GEMM at the UNet's Linear shapes (diagnostic for gemm_plain_bn)."""
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
L.rcdm_set_gemm_pair(0)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


SHAPES = [(970, 2048, 2048), (970, 6144, 2048), (970, 8192, 2048), (970, 2048, 8192)] if len(sys.argv) > 1 and sys.argv[1] == "prior" else None
if len(sys.argv) > 1 and sys.argv[1] == "small":
    SHAPES = [(640, 1280, 1280), (640, 3840, 1280), (640, 1280, 5120), (640, 1280, 2560)]
for (M, N, K) in SHAPES or [(10240, 640, 640), (10240, 1920, 640), (10240, 640, 2560), (640, 1280, 1280), (640, 3840, 1280), (640, 1280, 5120),
                  (2560, 1280, 1280), (2560, 3840, 1280), (2560, 1280, 5120), (40960, 320, 320), (40960, 960, 320), (40960, 320, 1280)]:
    a = torch.randn((M, K), device="cuda").to(dt)
    w = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(dt)
    b = torch.randn((N,), device="cuda")
    r = torch.randn((M, N), device="cuda").to(dt)
    out = torch.empty((M, N), dtype=dt, device="cuda")
    row = []
    for bn in (0, 64, 128, 160, 192):
        if bn == 160 and N % 160:
            row.append("   -  ")
            continue
        us = timeit(lambda: _lib.check(L.rcdm_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr(), out.data_ptr(), M, N, K, 0,
                                                   bn, 0, s)))
        row.append(f"{us:6.1f}")
    print(f"gemm M{M:6d} N{N:5d} K{K:5d} res1   auto {row[0]} | bn64 {row[1]} | bn128 {row[2]} | bn160 {row[3]} | bn192 {row[4]}", flush=True)
