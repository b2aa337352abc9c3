#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=10 --timeout=300 -k "parity" > gpurun_out/r2_pytest_ops_d.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_ops_d.log | cut -c1-600
timeout 200 python scripts/bench_ops.py flash 2>&1 | tee gpurun_out/r2_bench_flash_d.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -o gpurun_out/r2_gemm_small python scripts/prof_gemm_small.py > gpurun_out/r2_ncu_gemm_small.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_gemm_small.log
ls -la gpurun_out/*.ncu-rep
