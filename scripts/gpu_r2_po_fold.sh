#!/bin/bash
# round 2: proj_out folded over ff.net.2 (one two-segment GEMM) - whole-UNet parity + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/r2h_unet_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2h_unet_pytest.log)"
F="--steps 3 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline --no-pixels --no-extra-configs"
timeout 300 python bench.py $F > gpurun_out/r2h_bench_po_fold.log 2>&1; echo "fold rc=$?"; tail -1 gpurun_out/r2h_bench_po_fold.log | cut -c1-400
timeout 300 python bench.py $F --unet-option po_fold=0 > gpurun_out/r2h_bench_two_gemm.log 2>&1; echo "two-gemm rc=$?"; tail -1 gpurun_out/r2h_bench_two_gemm.log | cut -c1-400
