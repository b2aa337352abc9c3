#!/bin/bash
# round 2: prior - proj_out folded over ff.net.2 (rcdm_fold_proj + rcdm_gemm_cat): parity + A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prior_gpu.py -x -q -m gpu > gpurun_out/r2i_prior_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2i_prior_pytest.log)"
timeout 300 python bench.py --workload prior --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_prior_bench_fold.log 2>&1; echo "fold rc=$?"; tail -1 gpurun_out/r2i_prior_bench_fold.log | cut -c1-300
timeout 300 python bench.py --workload prior --steps 3 --warmup 3 --no-cpu-baseline --prior-two-gemm-proj-out > gpurun_out/r2i_prior_bench_two.log 2>&1; echo "two rc=$?"; tail -1 gpurun_out/r2i_prior_bench_two.log | cut -c1-300
