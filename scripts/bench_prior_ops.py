"""Warm per-kernel timings of one stage-1 prior layer at the shipped size (10 x 97 rows, width 2048), CUDA events
around 50 back-to-back launches per op, plus the roofline each op sits on (tensor: algorithmic FLOPs / time;
HBM-class ops: algorithmic bytes / time).  Weights of one layer only.  Output: one line per op + a JSON summary."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200 import _lib, ops  # noqa: E402

B, S, C, HEADS, MH = 10, 97, 2048, 32, 8
M = B * S
dt = torch.float16
g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dt)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


x = rnd(M, C)
res = rnd(M, C)
rows = []


def gemm_case(name, N, K, act=None, residual=False, count=1):
    a = rnd(M, K)
    w = rnd(N, K, scale=K ** -0.5)
    bias = torch.randn((N,), generator=g, device="cuda")
    out = torch.empty((M, N), dtype=dt, device="cuda")
    L = _lib.lib()
    flags = {None: 0, "gelu": 2}[act]
    r = res if residual else None

    def fn():
        _lib.check(L.rcdm_gemm_ex(1, a.data_ptr(), K, w.data_ptr(), bias.data_ptr(), r.data_ptr() if r is not None else None,
                                  0, out.data_ptr(), 0, M, N, K, flags, _lib.current_stream_ptr()))
    us = timeit(fn)
    fl = 2.0 * M * N * K
    rows.append(dict(op=name, us=us, count=count, tflops=fl / us / 1e6, bound="tensor"))


gemm_case("qkv / temporal qkv  N=6144 K=2048", 6144, 2048, count=3)
gemm_case("out-proj / proj_in / proj_out  N=2048 K=2048 +res", 2048, 2048, residual=True, count=5)
gemm_case("ff1 (GELU)  N=8192 K=2048", 8192, 2048, act="gelu")
gemm_case("ff2 / motion ff2  N=2048 K=8192 +res", 2048, 8192, residual=True, count=2)
# GEGLU (packed weights)
wg = rnd(16384, C, scale=C ** -0.5)
bg = torch.randn((16384,), generator=g, device="cuda")
us = timeit(lambda: ops.geglu_linear(x, wg, bg))  # includes the per-call pack (2 small kernels): upper bound
rows.append(dict(op="motion ff1 GEGLU  N=16384 K=2048 (incl. per-call pack)", us=us, count=1,
                 tflops=2.0 * M * 16384 * C / us / 1e6, bound="tensor"))
gamma = torch.ones((C,), device="cuda")
beta = torch.zeros((C,), device="cuda")
us = timeit(lambda: ops.layer_norm(x, gamma, beta))
rows.append(dict(op="layernorm 970 x 2048", us=us, count=6, gbs=2 * M * C * 2 / us / 1e3, bound="hbm"))
qkv = rnd(B, S, 3 * C)
kb = torch.zeros((B, S), device="cuda")
kb[:, 40:91] = -10000.0
us = timeit(lambda: ops.masked_attention(qkv, HEADS, kb, causal=True))
rows.append(dict(op="masked attention 320 heads x 97 x 97 x 64", us=us, count=1, gbs=4 * M * C * 2 / us / 1e3,
                 tflops=4.0 * B * S * S * C / us / 1e6, bound="hbm"))
qkv2 = rnd(M, 3 * C)
us = timeit(lambda: ops.temporal_attention(qkv2, B // 5, 5, S, MH))
rows.append(dict(op="temporal attention 2 x 97 x 8 heads x 5 x 5 x 256", us=us, count=2, gbs=4 * M * C * 2 / us / 1e3,
                 bound="hbm"))
tot = sum(r["us"] * r["count"] for r in rows)
for r in rows:
    extra = f"{r['tflops']:7.1f} TFLOP/s" if r["bound"] == "tensor" else f"{r['gbs']:7.1f} GB/s"
    print(f"{r['us']:8.1f} us x{r['count']}  {100 * r['us'] * r['count'] / tot:5.1f}%  {extra}  {r['op']}")
print(json.dumps(dict(layer_us=tot, layers=20, step_ms_from_ops=tot * 20 / 1e3, ops=rows)))
