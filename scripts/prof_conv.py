"""Deep-K 3x3 convolutions for `ncu --set full` (3 launches each, capture the last): 64x64 640->320 (K = 5760) and 32x32
640->640 (K = 5760), with the library's own scheduling (CTA pairs by heuristic)."""
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
for (n, h, cin, cout) in ((10, 64, 640, 320), (10, 32, 640, 640)):
    x = torch.randn((n, h, h, cin), device="cuda").to(dt)
    wp = (torch.randn((cout, 9 * cin), device="cuda") / math.sqrt(9 * cin)).to(dt)
    b = torch.randn((cout,), device="cuda")
    out = torch.empty((n, h, h, cout), dtype=dt, device="cuda")
    for _ in range(3):
        _lib.check(L.rcdm_conv3x3(1, x.data_ptr(), wp.data_ptr(), b.data_ptr(), None, out.data_ptr(), n, h, h, cin, cout, 1, 0, s))
    torch.cuda.synchronize()
