"""One launch of each short-key-range attention kernel (80 images x 4096 queries x 91 keys, d = 40) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rcdms_b200 import _lib, ops
L = _lib.lib()
b, h, sq, skv, d = 80, 8, 4096, 91, 40
q = torch.randn((b, sq, h * d), device="cuda").half()
k = torch.randn((b, skv, h * d), device="cuda").half()
v = torch.randn((b, skv, h * d), device="cuda").half()
for mode in (1, 0):
    prev = L.rcdm_debug_set_option(b"attn_short_kv", mode)
    for _ in range(3):
        ops.flash_attention(q, k, v, h)
    torch.cuda.synchronize()
    L.rcdm_debug_set_option(b"attn_short_kv", prev)
