#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/ffn_check.py time 2>&1 | tail -20 | tee gpurun_out/r2_ffn_check.log
