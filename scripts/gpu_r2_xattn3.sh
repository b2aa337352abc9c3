#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'cross_attn_mma|flash_attn4' -s 2 -c 1 -f -o gpurun_out/r2l_xattn_mma python scripts/prof_xattn.py > gpurun_out/r2l_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'flash_attn4' -s 2 -c 1 -f -o gpurun_out/r2l_xattn_flash python scripts/prof_xattn.py > gpurun_out/r2l_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/r2l_*
