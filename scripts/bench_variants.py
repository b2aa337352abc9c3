"""Times a few model GEMM shapes with the library selected by RCDM_LIB (bottleneck-analysis variants, see
scripts/build_variants.sh).  Numbers from variant builds are diagnostics, never bench values."""
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
tag = (os.environ.get("RCDM_LIB") or "x/product/x").split("/")[-2]


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


def gemm(M, N, K, res, geglu=0):
    a = torch.randn((M, K), device="cuda").to(dt)
    w = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(dt)
    b = torch.randn((N,), device="cuda")
    r = torch.randn((M, N), device="cuda").to(dt) if res else None
    out = torch.empty((M, N // 2 if geglu else N), dtype=dt, device="cuda")
    us = timeit(lambda: _lib.check(L.rcdm_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr() if res else None,
                                               out.data_ptr(), M, N, K, geglu, 0, 0, s)))
    print(f"{tag:8s} gemm M{M} N{N} K{K} res{int(res)} geglu{geglu}: {us:7.1f} us", flush=True)


def conv(n, h, cin, cout):
    x = torch.randn((n, h, h, cin), device="cuda").to(dt)
    wp = (torch.randn((cout, 9 * cin), device="cuda") / math.sqrt(9 * cin)).to(dt)
    b = torch.randn((cout,), device="cuda")
    out = torch.empty((n, h, h, cout), dtype=dt, device="cuda")
    us = timeit(lambda: _lib.check(L.rcdm_conv3x3(1, x.data_ptr(), wp.data_ptr(), b.data_ptr(), None, out.data_ptr(), n, h, h,
                                                  cin, cout, 1, 0, s)))
    print(f"{tag:8s} conv n{n} {h}x{h} {cin}->{cout}: {us:7.1f} us", flush=True)


for pair in (0, 2):
    L.rcdm_set_gemm_pair(pair)
    tag = tag.split("+")[0] + ("+pair" if pair else "")
    gemm(40960, 320, 320, True)
    gemm(40960, 320, 320, False)
    gemm(20480, 320, 320, False)
    gemm(10240, 320, 320, False)
    gemm(81920, 320, 320, False)
    gemm(40960, 960, 320, False)
    gemm(40960, 320, 1280, True)
    gemm(10240, 640, 640, True)
    gemm(10240, 1920, 640, False)
    gemm(2560, 1280, 1280, True)
    gemm(2560, 3840, 1280, False)
    gemm(40960, 2560, 320, False, 1)
    gemm(10240, 5120, 640, False, 1)
    gemm(2560, 10240, 1280, False, 1)
    conv(10, 64, 320, 320)
    conv(10, 32, 640, 640)
