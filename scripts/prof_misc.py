"""Launches for one `ncu --set full` capture of the non-GEMM kernels at the 512x512 shapes:
gn_fused (5-D stats, 64x64 C320 and per-frame), temporal tile kernel, flash attention d=40, LN-folded GEGLU GEMM."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200 import _lib, ops  # noqa: E402

dt = torch.float16
x = torch.randn((40960, 320), device="cuda").to(dt)
g = torch.ones((320,), device="cuda")
for _ in range(2):
    ops.group_norm(x, g, g, 32, 20480, 1e-5, True)
for _ in range(2):
    ops.group_norm(x, g, g, 32, 4096, 1e-6, False)
qkv = torch.randn((40960, 960), device="cuda").to(dt)
for _ in range(2):
    ops.temporal_attention(qkv, 2, 5, 4096, 8)
q = torch.randn((10, 4096, 320), device="cuda").to(dt)
for _ in range(2):
    ops.flash_attention(q, q, q, 8)
w = (torch.randn((2560, 320), device="cuda") / math.sqrt(320)).to(dt)
b = torch.randn((2560,), device="cuda")
for _ in range(2):
    ops.linear_ln(x, w, g, g, b, None, 1, True)
torch.cuda.synchronize()
