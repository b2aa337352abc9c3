#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_gemm_variants_f.log
timeout 200 python scripts/bench_variants.py >> gpurun_out/r2_gemm_variants_f.log 2>&1
RCDM_LIB=$PWD/rcdms_b200/_Cx10/librcdm_b200.so timeout 200 python scripts/bench_variants.py >> gpurun_out/r2_gemm_variants_f.log 2>&1
grep -v "+pair" gpurun_out/r2_gemm_variants_f.log
timeout 200 python scripts/bench_ops.py flash 2>&1 | tee gpurun_out/r2_bench_flash_f.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:flash_attn4 -c 3 -o gpurun_out/r2_flash python scripts/prof_flash.py > gpurun_out/r2_ncu_flash.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_ncu_flash.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs --no-eager-gpu-baseline > gpurun_out/r2_bench_f.log 2>&1
echo "bench rc=$?"; tail -c 1200 gpurun_out/r2_bench_f.log
