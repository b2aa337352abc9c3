"""bench.py contract on the CPU tier: the reference arms (the only legs that run without a GPU) print ONE JSON line with the
keys the driver reads, and our own arm fails loudly — never falls back — when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"}


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_prior_reference_arm_prints_one_json_line():
    r = _run("--workload", "prior", "--impl", "reference")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "frame-embeddings/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == os.cpu_count()
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_our_arm_fails_loudly_without_a_gpu():
    for extra in ([], ["--workload", "prior"]):
        r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", *extra, timeout=300)
        assert r.returncode != 0, "bench.py must not fall back to a CPU path"
        assert not [l for l in r.stdout.splitlines() if l.startswith("{")], r.stdout[-500:]


def test_stage2_reference_arm_line_and_config_keys():
    """The stage-2 reference arm at a small latent size: one JSON line, the SAME config keys / values our arm prints for
    the same flags (so the driver's same_config check holds), and ms_per_step = the measured duration of one bounded
    sample (so steps x ms_per_step fits in the run's wall time)."""
    import time
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    t0 = time.time()
    r = _run("--impl", "reference", "--latent", "8", "--ddim-steps", "10", "--steps", "2", "--warmup", "1")
    wall = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["steps"] == 2
    a = argparse.Namespace(latent=8, ddim_steps=10, dtype="fp16", clips=1, guidance=2.0, ctx_len=85)
    assert d["config"] == bench.config_dict(a, 1)
    assert d["steps"] * d["ms_per_step"] / 1e3 <= wall
    assert abs(d["value"] - 5.0 / (d["ms_per_step"] / 1e3 * 10)) < 1e-9 * max(1.0, d["value"])
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
