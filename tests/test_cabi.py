"""CPU-side checks of the C ABI: the shared library loads without a GPU, exports every symbol include/rcdm.h
declares (and the ctypes table binds exactly those), mirrors the reference's state-dict surface, and refuses to
compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import json
import os
import re

import pytest
import torch

from rcdms_b200 import _lib
from rcdms_b200.unet_spec import full_config, state_dict_spec, tiny_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from rcdms_b200.build import build
    build()
    return _lib.lib()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rcdm.h")).read()
    return sorted(set(re.findall(r"RCDM_API[^;(]*?\b(rcdm_\w+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/rcdm.h but not exported"
    assert sorted(_lib.SIGNATURES) == decl


def test_no_torch_or_cuda_runtime_dependency():
    """plain C ABI: the .so must not link libtorch / libc10 (torch types never cross the boundary)"""
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    names = [ln.split()[0] for ln in out.splitlines() if ln.strip()]  # library names only (load addresses are random hex)
    assert not [n for n in names if "torch" in n or "libc10" in n or "libcuda.so" in n], out


def _create(lib, cfg):
    from rcdms_b200.models.unet import _c_config
    h = C.c_void_p()
    cc = _c_config(cfg, torch.float16)
    _lib.check(lib.rcdm_unet_create(C.byref(cc), C.byref(h)))
    return h


def test_state_dict_surface_matches_reference(lib):
    h = _create(lib, full_config())
    n = lib.rcdm_unet_num_weights(h)
    buf = C.create_string_buffer(256)
    dims = (C.c_int64 * 4)()
    nd = C.c_int()
    got = []
    for i in range(n):
        _lib.check(lib.rcdm_unet_weight_info(h, i, buf, 256, dims, C.byref(nd)))
        got.append([buf.value.decode(), list(dims)[: nd.value]])
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")))
    assert sorted(got) == sorted(ref)
    assert lib.rcdm_unet_weights_missing(h) == 1286
    lib.rcdm_unet_destroy(h)


def test_python_module_mirrors_reference_state_dict():
    from rcdms_b200.models import UNet3DConditionModel
    with torch.device("meta"):
        m = UNet3DConditionModel.from_config(full_config())
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")))
    assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == ref
    assert m.config.sample_size == 64 and m.config.in_channels == 9
    with pytest.raises(TypeError):
        UNet3DConditionModel(bogus_key=1)
    from rcdms_b200.models.unet import _c_config
    with pytest.raises(NotImplementedError):
        _c_config(full_config(use_linear_projection=True), torch.float16)
    with pytest.raises(NotImplementedError):
        _c_config(full_config(unet_use_temporal_attention=True), torch.float16)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(lib):
    assert lib.rcdm_device_count() == 0
    h = _create(lib, tiny_config())
    assert lib.rcdm_unet_prepare(h, 2, 5, 8, 8, 7) != 0
    assert b"no CPU fallback" in lib.rcdm_last_error()
    assert lib.rcdm_gemm(1, None, None, None, None, None, 128, 64, 64, 0, 0, 0, None) != 0
    lib.rcdm_unet_destroy(h)
    from rcdms_b200.models import UNet3DConditionModel
    m = UNet3DConditionModel.from_config(tiny_config()).half()
    x = torch.zeros((2, 9, 5, 8, 8), dtype=torch.float16)
    with pytest.raises(RuntimeError):
        m(x, 1, encoder_hidden_states=torch.zeros((10, 7, 96), dtype=torch.float16))


def test_bad_config_is_rejected(lib):
    from rcdms_b200.models.unet import _c_config
    cc = _c_config(tiny_config(), torch.float16)
    cc.block_out_channels[0] = 100  # not a multiple of 64
    h = C.c_void_p()
    assert lib.rcdm_unet_create(C.byref(cc), C.byref(h)) != 0
    assert b"multiples of 64" in lib.rcdm_last_error()
