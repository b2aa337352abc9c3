#!/bin/bash
# round 2, call r: bottleneck experiments on the current GEMM (3 = A loaded once per tile, 5 = no epilogue, 6 = no stores /
# residual, 7 = B loaded once per CTA) + per-op forward profile of the product
mkdir -p gpurun_out
: > gpurun_out/r2_gemm_variants_r.log
for v in "" _Cx3 _Cx5 _Cx6 _Cx7; do
  if [ -n "$v" ]; then export RCDM_LIB=$PWD/rcdms_b200/$v/librcdm_b200.so; fi
  timeout 200 python scripts/bench_variants.py 2>&1 | grep -v "+pair" | tee -a gpurun_out/r2_gemm_variants_r.log
done
unset RCDM_LIB
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/r2_profile64_r.log 2>&1
echo "profile rc=$?"; head -40 gpurun_out/r2_profile64_r.log
