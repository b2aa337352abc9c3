#!/bin/bash
# Round-2 GPU call B: store-warp epilogue + GroupNorm statistics from the GEMM epilogue.  Every command under `timeout`.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout=300 > gpurun_out/r2_pytest_gpu_b.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/r2_pytest_gpu_b.log | cut -c1-900
timeout 200 python scripts/bench_variants.py > gpurun_out/r2_gemm_variants_b.log 2>&1; cat gpurun_out/r2_gemm_variants_b.log
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/r2_profile64_b.log 2>&1
echo "profile rc=$?"; head -24 gpurun_out/r2_profile64_b.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.log 2>&1
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench_b.log
