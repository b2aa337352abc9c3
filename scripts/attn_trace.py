"""Timeline of the 64x64-latent self-attention from a -DRCDM_ATTN_TRACE=1 variant build (RCDM_LIB=.../_Cxtrace/...):
per (CTA, K/V tile) clock64 stamps of the barrier hand-offs (attention.cuh: ATTN_STAMP).  Diagnostic only."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rcdms_b200 import _lib  # noqa: E402

dt = torch.float16
L = _lib.lib()
b, h, sq, skv, d = 10, 8, 4096, 4096, 40
q = torch.randn((b, sq, h * d), device="cuda").to(dt)
kv = torch.randn((b, skv, 2 * h * d), device="cuda").to(dt)
out = torch.empty_like(q)
c = h * d
for _ in range(3):
    _lib.check(L.rcdm_flash_attn(1, q.data_ptr(), c, kv.data_ptr(), kv.data_ptr() + c * 2, 2 * c, out.data_ptr(), c, b, h, sq, skv,
                                 d, 0, _lib.current_stream_ptr()))
torch.cuda.synchronize()
NC, NT = 2560, 64
st = np.zeros((NC, NT, 16), dtype=np.int64)
sm = np.zeros((NC,), dtype=np.int32)
raw = C.CDLL(os.environ["RCDM_LIB"])
raw.rcdm_debug_attn_trace_read.argtypes = [C.c_void_p, C.c_void_p]
assert raw.rcdm_debug_attn_trace_read(st.ctypes.data, sm.ctypes.data) == 0
np.save(os.path.join(ROOT, "gpurun_out", "r2_attn_trace.npy"), st[:600])
np.save(os.path.join(ROOT, "gpurun_out", "r2_attn_smid.npy"), sm)
# steady-state tiles 8..56 of CTAs that ran in the middle of the kernel
t0 = st[:, 0, 0].astype(np.float64)
order = np.argsort(t0)
mid = order[len(order) // 3: 2 * len(order) // 3]
S = st[mid][:, 8:56, :].astype(np.float64)
wait = S[:, :, 4:8] - S[:, :, 0:4]              # per warp: time blocked on s_full
busy = S[:, :, 8:12] - S[:, :, 4:8]             # per warp: S load + exponentials + P store
period = S[:, 1:, 8:12] - S[:, :-1, 8:12]       # per warp: p_full arrive to p_full arrive
print("per softmax warp 0..3 (clocks, mean over middle CTAs, tiles 8..55):")
print("  wait on s_full ", wait.mean(axis=(0, 1)).round(0))
print("  busy           ", busy.mean(axis=(0, 1)).round(0))
print("  tile period    ", period.mean(axis=(0, 1)).round(0))
last_arrive = S[:, :, 8:12].max(axis=2)
first_arrive = S[:, :, 8:12].min(axis=2)
print("  spread of the 4 p_full arrivals of a tile:", (last_arrive - first_arrive).mean().round(0))
# MMA thread: s_free(j) seen -> QK_{j+1} issued; p_full(j) seen relative to the last arrival; PV issue time
print("MMA thread (clocks):")
print("  s_free(j) seen after the LAST warp began tile j (its s_full wait end):", (S[:, :, 12] - S[:, :, 4:8].max(axis=2)).mean().round(0))
print("  QK_{j+1} issue duration:", (S[:, :, 13] - S[:, :, 12]).mean().round(0))
print("  p_full(j) seen after the last arrival:", (S[:, :, 14] - last_arrive).mean().round(0))
print("  PV_j issue duration (incl. v_full wait):", (S[:, :, 15] - S[:, :, 14]).mean().round(0))
# when does S_{j+1} become visible to the warps, relative to QK_{j+1} issued (stamp 13 of tile j)?  a warp that waited
# sees it at its wait end of tile j+1
seen = S[:, 1:, 4:8].min(axis=2)   # earliest warp out of the s_full(j+1) wait
print("  S_{j+1} first seen by a warp after QK_{j+1} issued:", (seen - S[:, :-1, 13]).mean().round(0))
print("  S_{j+1} first seen after the last p_full(j) arrival:", (seen - last_arrive[:, :-1]).mean().round(0))
# one CTA in detail
k = mid[len(mid) // 2]
base = st[k, 8, 0]
print(f"CTA {k} on SM {sm[k]}: tiles 8..13 relative clocks [wait begin x4 | wait end x4 | p_full arrive x4 | sfree qk pfull pv]")
for j in range(8, 14):
    print("  ", j, (st[k, j] - base).tolist())
co = [int(x) for x in np.nonzero(sm == sm[k])[0]]
print("CTAs on the same SM:", co[:40])
