#!/bin/bash
# round 2: folded LayerNorm in the stage-1 prior - parity (prior GPU checks + the chain test) and A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prior_gpu.py -x -q -m gpu > gpurun_out/r2g_prior_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2g_prior_pytest.log)"
timeout 300 python bench.py --workload prior --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_prior_bench_folded.log 2>&1; echo "folded rc=$?"; tail -1 gpurun_out/r2g_prior_bench_folded.log | cut -c1-900
timeout 300 python bench.py --workload prior --steps 3 --warmup 3 --no-cpu-baseline --prior-standalone-ln > gpurun_out/r2g_prior_bench_standalone.log 2>&1; echo "standalone rc=$?"; tail -1 gpurun_out/r2g_prior_bench_standalone.log | cut -c1-900
