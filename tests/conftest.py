import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_sessionstart(session):
    """RCDM_TEST_OPTS="sk_min=1,temporal_wide_all=0": run the suite with explicit library debug options (read by the TEST
    HARNESS and applied through rcdm_debug_set_option - the library itself reads no environment variables)."""
    opts = os.environ.get("RCDM_TEST_OPTS", "")
    if not opts:
        return
    from rcdms_b200 import _lib
    for kv in opts.split(","):
        k, v = kv.split("=")
        if _lib.lib().rcdm_debug_set_option(k.strip().encode(), int(v)) < 0:
            raise pytest.UsageError(f"unknown library option {k}")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
