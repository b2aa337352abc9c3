"""CUDA-event timing of the model's GEMM / conv shapes through the C ABI with the stream-K decomposition off and on.
usage: python scripts/bench_gemm_shapes.py"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200 import _lib  # noqa: E402

dt = torch.float16
L = _lib.lib()


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
    return e0.elapsed_time(e1) / reps * 1e3  # us


def both(name, flops, run):
    res = []
    for pair, sk in ((0, 0), (0, 24), (0, 8), (2, 8)):
        L.rcdm_set_gemm_pair(pair)
        L.rcdm_set_stream_k_min(sk)
        res.append(timeit(run))
    L.rcdm_set_gemm_pair(1)
    L.rcdm_set_stream_k_min(24)
    tf = [flops / r / 1e6 for r in res]
    print(f"{name:46s} 1cta {res[0]:6.1f} us ({tf[0]:5.0f}) | 1cta+sk24 {res[1]:6.1f} ({tf[1]:5.0f}) | 1cta+sk8 {res[2]:6.1f} ({tf[2]:5.0f}) | "
          f"pair+sk8 {res[3]:6.1f} us ({tf[3]:5.0f} TF/s)", flush=True)


def gemm_case(M, N, K, res=True):
    a = torch.randn((M, K), device="cuda").to(dt)
    w = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(dt)
    b = torch.randn((N,), device="cuda")
    r = torch.randn((M, N), device="cuda").to(dt) if res else None
    out = torch.empty((M, N), dtype=dt, device="cuda")
    s = _lib.current_stream_ptr()

    def run():
        _lib.check(L.rcdm_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr() if res else None, out.data_ptr(),
                               M, N, K, 0, 0, 0, s))
    both(f"gemm M{M} N{N} K{K} res{int(res)}", 2.0 * M * N * K, run)


def conv_case(n, h, cin, cout, stride=1):
    x = torch.randn((n, h, h, cin), device="cuda").to(dt)
    wp = (torch.randn((cout, 9 * cin), device="cuda") / math.sqrt(9 * cin)).to(dt)
    b = torch.randn((cout,), device="cuda")
    out = torch.empty((n, h // stride, h // stride, cout), dtype=dt, device="cuda")
    s = _lib.current_stream_ptr()

    def run():
        _lib.check(L.rcdm_conv3x3(1, x.data_ptr(), wp.data_ptr(), b.data_ptr(), None, out.data_ptr(), n, h, h, cin, cout,
                                  stride, 0, s))
    M = n * (h // stride) ** 2
    both(f"conv3x3 n{n} {h}x{h} {cin}->{cout} s{stride} (M{M} K{9 * cin})", 2.0 * M * cout * 9 * cin, run)


if __name__ == "__main__":
    for (M, N, K) in [(40960, 320, 320), (40960, 320, 1280), (40960, 960, 320), (10240, 640, 640), (10240, 640, 2560),
                      (10240, 1920, 640), (2560, 1280, 1280), (2560, 1280, 5120), (2560, 3840, 1280), (640, 1280, 1280),
                      (640, 1280, 5120), (640, 3840, 1280)]:
        gemm_case(M, N, K)
    for (n, h, cin, cout) in [(10, 64, 320, 320), (10, 64, 640, 320), (10, 64, 960, 320), (10, 32, 640, 640),
                              (10, 32, 1280, 640), (10, 32, 1920, 640), (10, 16, 1280, 1280), (10, 16, 2560, 1280),
                              (10, 8, 1280, 1280), (10, 8, 2560, 1280)]:
        conv_case(n, h, cin, cout)
    conv_case(10, 64, 320, 320, 2)
    conv_case(10, 32, 640, 640, 2)
    gemm_case(8192, 8320, 8192, res=False)
