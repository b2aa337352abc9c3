#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:flash_attn4 -s 2 -c 1 -f -o gpurun_out/r2_flash_t python scripts/prof_flash.py > gpurun_out/r2_ncu_flash_t.log 2>&1
echo "ncu rc=$?"
