#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/ffn_check.py time 2>&1 | tail -9 | cut -c1-200 | tee gpurun_out/r2_ffn_check.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/r2_pytest_gpu_ffn.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_ffn.log | cut -c1-300
timeout 300 python - <<'PY' 2>&1 | tail -4
import json, subprocess, sys
# the fused path inside the full forward: profile with the option on
sys.path.insert(0, '.')
from rcdms_b200 import _lib
L = _lib.lib()
L.rcdm_debug_set_option(b"ffn_fused", 1)
import runpy
sys.argv = ['profile_forward.py', '64', '1']
runpy.run_path('scripts/profile_forward.py', run_name='__main__')
PY
