#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/ffn_check.py time 2>&1 | tail -10 | cut -c1-200 | tee gpurun_out/r2_ffn_check.log
RCDM_LIB=$PWD/rcdms_b200/_Cxtrace/librcdm_b200.so timeout 120 python scripts/ffn_trace.py 2>&1 | tail -8 | tee gpurun_out/r2_ffn_trace.log
