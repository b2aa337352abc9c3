"""Two full-size UNet forwards (first = warm-up) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcdms_b200.models import UNet3DConditionModel  # noqa: E402
from rcdms_b200.synthetic import synthetic_state_dict  # noqa: E402
from rcdms_b200.unet_spec import full_config  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = full_config()
m = UNet3DConditionModel.from_config(cfg)
m.load_state_dict(synthetic_state_dict(cfg, seed=0))
m = m.cuda().half()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((2, 9, 5, 64, 64), device="cuda", generator=g).half()
ctx = torch.randn((10, 85, 768), device="cuda", generator=g).half()
for _ in range(n):
    y = m(x, 501, encoder_hidden_states=ctx)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
