"""Per-op (CUDA-event) profile of one full-size UNet forward, grouped by (kind, shape).
usage: python scripts/profile_forward.py [latent=64] [clips=1]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200.models import UNet3DConditionModel  # noqa: E402
from rcdms_b200.synthetic import synthetic_state_dict  # noqa: E402
from rcdms_b200.unet_spec import full_config  # noqa: E402

latent = int(sys.argv[1]) if len(sys.argv) > 1 else 64
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = full_config()
m = UNet3DConditionModel.from_config(cfg)
m.load_state_dict(synthetic_state_dict(cfg, seed=0))
m = m.cuda().half()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((2 * clips, 9, 5, latent, latent), device="cuda", generator=g).half()
ctx = torch.randn((10 * clips, 85, 768), device="cuda", generator=g).half()
prof = m.profile(x, 501.0, ctx, reps=5)
agg = {}
for p in prof:
    key = (p["kind"],) + tuple(p["shape"])
    a = agg.setdefault(key, dict(ms=0.0, n=0, flops=0.0, bytes=0.0))
    a["ms"] += p["ms"]
    a["n"] += 1
    a["flops"] += p["flops"]
    a["bytes"] += p["bytes"]
tot = sum(a["ms"] for a in agg.values())
print(f"total {tot:.3f} ms over {len(prof)} ops")
rows = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
out = []
for k, a in rows:
    tf = a["flops"] / max(a["ms"], 1e-9) / 1e9
    gb = a["bytes"] / max(a["ms"], 1e-9) / 1e6
    print(f"{k[0]:14s} {str(k[1:]):26s} n={a['n']:3d} ms={a['ms']:8.3f} ({100 * a['ms'] / tot:4.1f}%) us/op={1e3 * a['ms'] / a['n']:8.1f} "
          f"TF/s={tf:7.1f} GB/s={gb:7.1f}")
    out.append(dict(kind=k[0], shape=k[1:], **a))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(total_ms=tot, rows=out), open(os.path.join(ROOT, "gpurun_out", f"forward_profile_{latent}.json"), "w"), indent=1)
