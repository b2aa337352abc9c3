"""Short-key-range attention (cross-attention to 85 / 91 context tokens, 8x8 self-attention): register-resident mma.sync
kernel (library option attn_short_kv = 1, default) against the tcgen05 flash kernel (0) on the UNet's shapes; us per launch,
CUDA events over 40 launches rotating three input sets (> L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcdms_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
shapes = [(10, 8, 4096, 85, 40), (10, 8, 1024, 85, 80), (10, 8, 256, 85, 160), (10, 8, 64, 85, 160), (10, 8, 64, 64, 160),
          (80, 8, 4096, 91, 40), (80, 8, 1024, 91, 80), (80, 8, 256, 91, 160)]
g = torch.Generator(device="cuda").manual_seed(0)
for (b, h, sq, skv, d) in shapes:
    sets = [(torch.randn((b, sq, h * d), generator=g, device="cuda").half(),
             torch.randn((b, skv, h * d), generator=g, device="cuda").half(),
             torch.randn((b, skv, h * d), generator=g, device="cuda").half()) for _ in range(3)]
    res = {}
    for mode in (1, 0):
        prev = L.rcdm_debug_set_option(b"attn_short_kv", mode)
        for i in range(6):
            ops.flash_attention(*sets[i % 3], h)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(42):
            ops.flash_attention(*sets[i % 3], h)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / 42 * 1e3
        L.rcdm_debug_set_option(b"attn_short_kv", prev)
    byt = 2.0 * b * sq * h * d * 2
    print(f"attn b{b} h{h} Sq{sq} Skv{skv} d{d}: mma.sync {res[1]:7.1f} us ({byt / res[1] / 1e6:6.2f} TB/s of q + out) | "
          f"flash {res[0]:7.1f} us", flush=True)
