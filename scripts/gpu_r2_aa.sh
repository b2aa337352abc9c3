#!/bin/bash
mkdir -p gpurun_out
RCDM_LIB=$PWD/rcdms_b200/_Cxtrace/librcdm_b200.so timeout 300 python scripts/gemm_trace.py 2>&1 | tee gpurun_out/r2_gemm_trace.log
