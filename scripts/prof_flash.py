import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rcdms_b200 import _lib
dt = torch.float16
L = _lib.lib()
b, h, sq, skv, d = 10, 8, 4096, 4096, 40
q = torch.randn((b, sq, h * d), device="cuda").to(dt)
kv = torch.randn((b, skv, 2 * h * d), device="cuda").to(dt)
out = torch.empty_like(q)
c = h * d
for _ in range(3):
    _lib.check(L.rcdm_flash_attn(1, q.data_ptr(), c, kv.data_ptr(), kv.data_ptr() + c * 2, 2 * c, out.data_ptr(), c, b, h, sq, skv, d, 0,
                                 _lib.current_stream_ptr()))
torch.cuda.synchronize()
