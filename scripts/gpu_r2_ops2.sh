#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q --maxfail=5 --timeout=600 > gpurun_out/r2_pytest_ops_last2.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_ops_last2.log | cut -c1-300
