"""Timeline of the fused GEGLU feed-forward kernel from a -DRCDM_FFN_TRACE=1 variant build (CTA 7, first 128 chunks)."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rcdms_b200 import ops  # noqa: E402

M, Cc, J = 40960, 320, 1280
dt = torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
y = torch.randn((M, Cc), device="cuda", generator=g).to(dt)
w1 = (torch.randn((2 * J, Cc), device="cuda", generator=g) / math.sqrt(Cc)).to(dt)
b1 = torch.randn((2 * J,), device="cuda", generator=g)
w2 = (torch.randn((Cc, J), device="cuda", generator=g) / math.sqrt(J)).to(dt)
b2 = torch.randn((Cc,), device="cuda", generator=g)
for _ in range(3):
    ops.ffn_geglu_ln(y, w1, b1, torch.ones((Cc,), device="cuda"), torch.zeros((Cc,), device="cuda"), w2, b2)
torch.cuda.synchronize()
raw = C.CDLL(os.environ["RCDM_LIB"])
raw.rcdm_debug_ffn_trace_read.argtypes = [C.c_void_p]
st = np.zeros((128, 16), dtype=np.int64)
assert raw.rcdm_debug_ffn_trace_read(st.ctypes.data) == 0
names = ["P0 b1 slot free", "P0 b1 issued", "P1 b2 slot free", "P1 b2 issued", "MMA g1 start", "MMA acc1 free", "MMA b1 landed",
         "MMA g1 issued", "MMA g2 start", "MMA a2 full", "MMA b2 landed", "EPI wait begin", "EPI acc1 full", "EPI math done",
         "EPI a2 free", "EPI a2 written"]
base = st[8, 4]
print("chunk " + " ".join(f"{n[:9]:>9s}" for n in names))
for c in range(8, 20):
    print(f"{c:5d} " + " ".join(f"{int(v - base):9d}" for v in st[c]))
S = st[10:38].astype(np.float64)
per = (S[-1] - S[0]) / (len(S) - 1)
print("mean period per chunk by stamp:", {names[i]: int(per[i]) for i in range(16)})
d = lambda a, b: float(np.mean(S[:, a] - S[:, b]))
print(f"P0: issue cost {d(1, 0):.0f} | P1: issue cost (2 boxes) {d(3, 2):.0f}")
print(f"MMA: wait acc1 free {d(5, 4):.0f} | wait b1 landed {d(6, 5):.0f} | issue GEMM1 {d(7, 6):.0f} | "
      f"gemm2: wait a2 {d(9, 8):.0f} | wait b2 {d(10, 9):.0f}")
print(f"EPI: wait acc1 full {d(12, 11):.0f} | ld + math {d(13, 12):.0f} | wait a2 free {d(14, 13):.0f} | write + fence {d(15, 14):.0f}")
print(f"b1: issued -> landed (MMA sees) {d(6, 1):.0f} | b2: issued -> landed {float(np.mean(S[:, 10] - S[:, 3])):.0f}")
