#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'cross_attn_mma' -s 2 -c 1 -f -o gpurun_out/r2o_xattn_v3 python scripts/prof_xattn.py > gpurun_out/r2o_ncu.log 2>&1; echo "ncu rc=$?"
