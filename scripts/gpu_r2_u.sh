#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_bench_flash_u.log
for v in "" _Cxs0 _Cxs32 _Cxs300; do
  if [ -n "$v" ]; then export RCDM_LIB=$PWD/rcdms_b200/$v/librcdm_b200.so; fi
  echo "--- ${v:-product (100 ns)}" | tee -a gpurun_out/r2_bench_flash_u.log
  timeout 200 python scripts/bench_ops.py flash 2>&1 | tee -a gpurun_out/r2_bench_flash_u.log
done
