#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --timeout=300 > gpurun_out/r2_pytest_gpu_i.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/r2_pytest_gpu_i.log | cut -c1-700
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/r2_profile64_i.log 2>&1
echo "profile rc=$?"; head -3 gpurun_out/r2_profile64_i.log; grep "conv3x3" gpurun_out/r2_profile64_i.log | head -12
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs --no-eager-gpu-baseline > gpurun_out/r2_bench_i.log 2>&1
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_i.log | cut -c1-700
