#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=5 --timeout=300 > gpurun_out/r2_pytest_ops_x.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/r2_pytest_ops_x.log | cut -c1-300
timeout 200 python scripts/bench_ops.py flash 2>&1 | tee gpurun_out/r2_bench_flash_x.log
RCDM_LIB=$PWD/rcdms_b200/_Cxtrace/librcdm_b200.so timeout 300 python scripts/attn_trace.py 2>&1 | tee gpurun_out/r2_attn_trace_x.log
