#!/bin/bash
mkdir -p gpurun_out
RCDM_LIB=$PWD/rcdms_b200/_Cxtrace/librcdm_b200.so timeout 120 python scripts/ffn_trace.py 2>&1 | tail -30 | tee gpurun_out/r2_ffn_trace.log
