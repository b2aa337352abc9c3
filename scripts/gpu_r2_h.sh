#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py -m gpu -q --maxfail=10 --timeout=300 > gpurun_out/r2_pytest_vae.log 2>&1
echo "pytest vae rc=$?"; tail -40 gpurun_out/r2_pytest_vae.log | cut -c1-400
timeout 300 python scripts/bench_gemm_shapes.py > gpurun_out/r2_gemm_shapes_h.log 2>&1; cat gpurun_out/r2_gemm_shapes_h.log | cut -c1-200
