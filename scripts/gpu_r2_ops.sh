#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=5 --timeout=300 > gpurun_out/r2_pytest_ops_last.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/r2_pytest_ops_last.log | cut -c1-300
