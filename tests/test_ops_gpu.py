"""GPU parity of every CUDA kernel, called through the C ABI (ctypes), vs an fp32 torch reference fed the same
16-bit-rounded inputs.  Tolerances (written here): fp16 rtol 1e-3 / atol 1e-3*rms(ref) per op
(x2 for norm/GEGLU epilogues, x3 for attention, which chain two roundings); bf16 rtol 1.6e-2."""
import pytest
import torch

import parity_checks as pc

pytestmark = pytest.mark.gpu

_CHECKS = list(pc.all_op_checks()) if torch.cuda.is_available() else []


@pytest.mark.parametrize("idx", range(len(_CHECKS)))
def test_op_parity(idx):
    r = _CHECKS[idx]()
    assert r["ok"], r


def test_native_library_is_loaded_and_counts_launches():
    from rcdms_b200 import _lib
    L = _lib.lib()
    before = L.rcdm_kernel_launches()
    pc.check_linear(256, 160, 64, torch.float16)
    assert L.rcdm_kernel_launches() > before
    assert L.rcdm_device_count() >= 1


def test_errors_are_python_exceptions():
    from rcdms_b200 import _lib, ops
    a = torch.zeros((8, 12), dtype=torch.float16, device="cuda")   # K = 12 is not a multiple of 8
    w = torch.zeros((16, 12), dtype=torch.float16, device="cuda")
    with pytest.raises(_lib.RcdmError):
        ops.linear(a, w)
    with pytest.raises(TypeError):
        ops.linear(a.float(), w.float())
