"""Timeline of the K = 320 GEMMs from a -DRCDM_GEMM_TRACE=1 variant build (RCDM_LIB=.../_Cxtrace/...): clock64 stamps of the
hand-offs between producer / MMA / epilogue / store roles for the first 16 tiles of every CTA.  Diagnostic only."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rcdms_b200 import _lib  # noqa: E402

dt = torch.float16
L = _lib.lib()
raw = C.CDLL(os.environ["RCDM_LIB"])
raw.rcdm_debug_gemm_trace_read.argtypes = [C.c_void_p]
names = ["P0 slot free (kb0)", "P1 slot free (kb1)", "P0 last issued", "P1 last issued", "MMA acc free", "MMA first kb landed",
         "MMA last kb landed", "MMA committed", "EPI wait acc begin", "EPI acc full", "EPI staging ready", "EPI acc released",
         "ST stg_full seen", "ST stores issued", "ST store read done", "ST residual issued / freed"]


def run(M, N, K, res, geglu=0):
    a = torch.randn((M, K), device="cuda").to(dt)
    w = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(dt)
    b = torch.randn((N,), device="cuda")
    r = torch.randn((M, N), device="cuda").to(dt) if res else None
    out = torch.empty((M, N // 2 if geglu else N), dtype=dt, device="cuda")
    for _ in range(3):
        _lib.check(L.rcdm_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr() if res else None, out.data_ptr(), M, N, K,
                               geglu, 0, 0, _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    st = np.zeros((148, 16, 16), dtype=np.int64)
    assert raw.rcdm_debug_gemm_trace_read(st.ctypes.data) == 0
    print(f"=== gemm M{M} N{N} K{K} res{int(res)} geglu{geglu}")
    cta = 40
    base = st[cta, 0, 0]
    ntile = int((st[cta, :, 9] > 0).sum())
    print(f"CTA {cta}: {ntile} traced tiles; clocks relative to its first producer stamp")
    print("tile " + " ".join(f"{i:>6d}" for i in range(16)))
    for t in range(min(ntile, 8)):
        print(f"{t:4d} " + " ".join(f"{int(v - base) if v else -1:6d}" for v in st[cta, t]))
    # steady-state tile period of every role (tiles 1..ntile-2), mean over CTAs
    per = []
    for c in range(148):
        n = int((st[c, :, 9] > 0).sum())
        if n >= 4:
            per.append([(st[c, n - 2, s] - st[c, 1, s]) / (n - 3) if st[c, n - 2, s] and st[c, 1, s] else np.nan for s in range(16)])
    per = np.nanmean(np.array(per, dtype=np.float64), axis=0)
    print("mean tile period by stamp:", {names[i]: int(per[i]) for i in range(16) if per[i] == per[i]})
    S = st[:, 1:3, :].astype(np.float64)
    ok = S[:, :, 9] > 0
    def d(a_, b_):
        return float(np.mean((S[:, :, a_] - S[:, :, b_])[ok]))
    print(f"  EPI waits for the accumulator {d(9, 8):.0f} | then for staging {d(10, 9):.0f} | column loop {d(11, 10):.0f} | "
          f"store thread sees stg_full {d(12, 11):.0f} later | issues stores in {d(13, 12):.0f} | store read {d(14, 13):.0f} | "
          f"residual / free {d(15, 14):.0f}")
    print(f"  MMA: acc free -> first k-block landed {d(5, 4):.0f} | first -> last k-block landed {d(6, 5):.0f} | -> committed {d(7, 6):.0f}")
    print(f"  producers: slot free for the tile's first k-block -> last k-block issued: P0 {d(2, 0):.0f}  P1 {d(3, 1):.0f}")


run(40960, 320, 320, True)
run(40960, 960, 320, False)
run(40960, 2560, 320, False, 1)
run(10240, 640, 640, True)
