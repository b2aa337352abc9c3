"""Launch each hot kernel once at its real 512x512 shape (for `ncu -k regex:...`), or time them with CUDA events.
usage: python scripts/prof_kernels.py [time]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import parity_checks as pc  # noqa: E402

dt = torch.float16
CASES = {
    "flash_64": lambda: pc.check_flash(10, 8, 4096, 4096, 40, dt),
    "flash_32": lambda: pc.check_flash(10, 8, 1024, 1024, 80, dt),
    "flash_cross": lambda: pc.check_flash(10, 8, 4096, 85, 40, dt),
    "gemm_320_res": lambda: pc.check_linear(40960, 320, 320, dt, residual=True),
    "gemm_qkv": lambda: pc.check_linear(40960, 960, 320, dt, bias=False),
    "gemm_ff2": lambda: pc.check_linear(40960, 320, 1280, dt, residual=True),
    "gemm_1280": lambda: pc.check_linear(2560, 1280, 1280, dt, residual=True),
    "gemm_640": lambda: pc.check_linear(10240, 640, 640, dt, residual=True),
    "geglu_320": lambda: pc.check_geglu(40960, 320, dt),
    "conv_320": lambda: pc.check_conv3x3(10, 64, 64, 320, 320, 1, dt),
    "conv_1280": lambda: pc.check_conv3x3(10, 16, 16, 1280, 1280, 1, dt),
    "gn_320": lambda: pc.check_groupnorm(2, 5 * 4096, 320, dt),
    "ln_320": lambda: pc.check_layernorm(40960, 320, dt, pe=True),
    "temporal": lambda: pc.check_temporal(2, 5, 4096, 8, 40, dt),
}

if __name__ == "__main__":
    timing = len(sys.argv) > 1 and sys.argv[1] == "time"
    names = sys.argv[2:] if len(sys.argv) > 2 else list(CASES)
    for name in names:
        r = CASES[name]()
        torch.cuda.synchronize()
        print(name, "ok" if r["ok"] else "FAIL", round(r["max_abs"], 5), flush=True)
