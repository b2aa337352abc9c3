#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prior_gpu.py -m gpu -q --timeout=180 > gpurun_out/pytest_prior.log 2>&1
echo "pytest prior rc=$?"; tail -30 gpurun_out/pytest_prior.log | cut -c1-400
timeout 200 python scripts/bench_prior_ops.py > gpurun_out/bench_prior_ops.log 2>&1
echo "ops rc=$?"; head -12 gpurun_out/bench_prior_ops.log | cut -c1-200
timeout 300 python scripts/bench_prior.py --steps 100 --reps 3 > gpurun_out/bench_prior.log 2>&1
echo "bench prior rc=$?"; tail -1 gpurun_out/bench_prior.log | cut -c1-500
