#!/bin/bash
# Round-2 GPU call A: L2/TMA bandwidth ceiling, GEMM bottleneck variants, full GPU test suite (new full-shape parity
# tests included), per-op profile and a short bench line.  Every command under `timeout`.
mkdir -p gpurun_out
timeout 300 scripts/micro/l2_tma_bench.out > gpurun_out/r2_l2_tma_bench.log 2>&1
echo "l2 bench rc=$?"; cat gpurun_out/r2_l2_tma_bench.log
: > gpurun_out/r2_gemm_variants.log
timeout 200 python scripts/bench_variants.py >> gpurun_out/r2_gemm_variants.log 2>&1
for n in 4 5 6 7 9; do
  RCDM_LIB=$PWD/rcdms_b200/_Cx$n/librcdm_b200.so timeout 200 python scripts/bench_variants.py >> gpurun_out/r2_gemm_variants.log 2>&1
done
cat gpurun_out/r2_gemm_variants.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --timeout=300 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r2_pytest_gpu.log | cut -c1-800
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/r2_profile64_a.log 2>&1
echo "profile rc=$?"; head -70 gpurun_out/r2_profile64_a.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_a.log 2>&1
echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2_bench_a.log
