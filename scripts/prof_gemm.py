import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rcdms_b200 import _lib
dt = torch.float16
L = _lib.lib()
def run(M, N, K, bias, res, bn=0):
    a = torch.randn((M, K), device="cuda").to(dt)
    w = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(dt)
    b = torch.randn((N,), device="cuda") if bias else None
    r = torch.randn((M, N), device="cuda").to(dt) if res else None
    out = torch.empty((M, N), dtype=dt, device="cuda")
    for _ in range(3):
        _lib.check(L.rcdm_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr() if bias else None, r.data_ptr() if res else None,
                               out.data_ptr(), M, N, K, 0, bn, 0, _lib.current_stream_ptr()))
    torch.cuda.synchronize()
run(40960, 320, 320, 0, 0)
run(40960, 320, 320, 1, 1)
run(4096, 4160, 4096, 0, 0)
