"""CUDA-event timing of single kernels through the C ABI at the real 512x512 shapes (steady state, 20 reps)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200 import _lib, ops  # noqa: E402

dt = torch.float16


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


def gemm_case(M, N, K, bias, res, bn=0):
    a = torch.randn((M, K), device="cuda").to(dt)
    w = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(dt)
    b = torch.randn((N,), device="cuda") if bias else None
    r = torch.randn((M, N), device="cuda").to(dt) if res else None
    out = torch.empty((M, N), dtype=dt, device="cuda")
    L = _lib.lib()
    s = _lib.current_stream_ptr()

    def run():
        _lib.check(L.rcdm_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr() if bias else None,
                               r.data_ptr() if res else None, out.data_ptr(), M, N, K, 0, bn, 0, s))
    us = timeit(run)
    print(f"gemm M{M} N{N} K{K} bias{int(bias)} res{int(res)} bn{bn}: {us:8.1f} us  {2 * M * N * K / us / 1e6:7.1f} TF/s", flush=True)


def flash_case(b, h, sq, skv, d):
    q = torch.randn((b, sq, h * d), device="cuda").to(dt)
    kv = torch.randn((b, skv, 2 * h * d), device="cuda").to(dt)
    out = torch.empty_like(q)
    L = _lib.lib()
    s = _lib.current_stream_ptr()
    c = h * d

    def run():
        _lib.check(L.rcdm_flash_attn(1, q.data_ptr(), c, kv.data_ptr(), kv.data_ptr() + c * 2, 2 * c, out.data_ptr(), c,
                                     b, h, sq, skv, d, 0, s))
    us = timeit(run)
    print(f"flash b{b} h{h} Sq{sq} Skv{skv} d{d}: {us:8.1f} us  {4 * b * h * sq * skv * d / us / 1e6:7.1f} TF/s "
          f"({us * 1.965e3 * 148 / (b * h * ((sq + 127) // 128) * ((skv + 127) // 128)):.0f} cyc/tile/SM)", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "flash":
        flash_case(10, 8, 4096, 4096, 40)
        flash_case(10, 8, 1024, 1024, 80)
        flash_case(10, 8, 256, 256, 160)
        flash_case(10, 8, 4096, 85, 40)
        flash_case(10, 8, 1024, 85, 80)
        sys.exit(0)
    for bias, res in ((0, 0), (1, 0), (1, 1)):
        gemm_case(40960, 320, 320, bias, res)
    gemm_case(40960, 320, 320, 1, 1, bn=128)
    gemm_case(40960, 320, 320, 1, 1, bn=64)
    gemm_case(40960, 960, 320, 0, 0)
    gemm_case(40960, 320, 1280, 1, 1)
    gemm_case(10240, 640, 640, 1, 1)
    gemm_case(2560, 1280, 1280, 1, 1)
    gemm_case(2560, 1280, 1280, 1, 1, bn=128)
    gemm_case(2560, 1280, 1280, 1, 1, bn=64)
    gemm_case(640, 1280, 1280, 1, 1)
    gemm_case(8192, 8192, 8192, 0, 0, bn=128)
    gemm_case(8192, 8320, 8192, 0, 0)
    x = torch.randn((40960, 320), device="cuda").to(dt)
    g = torch.ones((320,), device="cuda")
    print("layernorm 40960x320: %.1f us" % timeit(lambda: ops.layer_norm(x, g, g)))
    print("groupnorm 2x20480x320: %.1f us" % timeit(lambda: ops.group_norm(x, g, g, 32, 20480, 1e-5, True)))
    print("groupnorm 10x4096x320: %.1f us" % timeit(lambda: ops.group_norm(x, g, g, 32, 4096, 1e-6, False)))
    qkv = torch.randn((40960, 960), device="cuda").to(dt)
    print("temporal 2x5x4096 C320: %.1f us" % timeit(lambda: ops.temporal_attention(qkv, 2, 5, 4096, 8)))
    y = torch.empty_like(x)
    print("torch copy 26MB: %.1f us" % timeit(lambda: y.copy_(x)))
