#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout=300 > gpurun_out/r2_pytest_gpu_c.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/r2_pytest_gpu_c.log | cut -c1-900
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/r2_profile64_c.log 2>&1
echo "profile rc=$?"; grep -i "groupnorm\|total" gpurun_out/r2_profile64_c.log
