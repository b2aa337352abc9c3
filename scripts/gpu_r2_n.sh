#!/bin/bash
# round 2, call n: TMA box-size microbenchmark; GEGLU tile width 160 in the product build (ops parity + bench)
mkdir -p gpurun_out
timeout 120 scripts/micro/tma_box_bench.out 2>&1 | tee gpurun_out/r2_tma_box_bench.log
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=5 --timeout=300 > gpurun_out/r2_pytest_ops_n.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/r2_pytest_ops_n.log | cut -c1-300
timeout 200 python scripts/bench_variants.py 2>&1 | grep -v "+pair" | tee gpurun_out/r2_gemm_variants_n.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_n.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n.log | cut -c1-700
