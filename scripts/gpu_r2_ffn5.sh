#!/bin/bash
mkdir -p gpurun_out
echo "--- pair kernel (ffn_fused = 2)"
FFN_MODE=2 timeout 90 python scripts/ffn_check.py time 2>&1 | tail -9 | cut -c1-200 | tee gpurun_out/r2_ffn_check_pair.log
