#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_gemm_variants_g.log
RCDM_LIB=$PWD/rcdms_b200/_Cx11/librcdm_b200.so timeout 200 python scripts/bench_variants.py >> gpurun_out/r2_gemm_variants_g.log 2>&1
grep -v "+pair" gpurun_out/r2_gemm_variants_g.log
