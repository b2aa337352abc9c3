#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/bench_variants.py 2>&1 | tee gpurun_out/r2_gemm_variants_ab.log
