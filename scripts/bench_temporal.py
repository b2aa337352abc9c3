"""Times the temporal attention kernel at the three UNet levels (diagnostic; RCDM_LIB selects a variant build)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200 import ops  # noqa: E402

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
    return e0.elapsed_time(e1) / reps * 1e3


tag = (os.environ.get("RCDM_LIB") or "x/product/x").split("/")[-2]
for hw, C in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
    rows = 2 * 5 * hw
    qkv = torch.randn((rows, 3 * C), device="cuda").to(dt)
    us = timeit(lambda: ops.temporal_attention(qkv, 2, 5, hw, 8))
    print(f"{tag:8s} temporal 2x5x{hw} C{C}: {us:6.1f} us  {rows * C * 4 * 2 / us / 1e6:5.2f} TB/s")
