#!/bin/bash
# round 2, call o: two TMA producer threads in the GEMM (product) against 1 / 3 producers; ops parity; bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=5 --timeout=300 > gpurun_out/r2_pytest_ops_o.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/r2_pytest_ops_o.log | cut -c1-300
for v in "" _Cxp1 _Cxp3; do
  if [ -n "$v" ]; then export RCDM_LIB=$PWD/rcdms_b200/$v/librcdm_b200.so; fi
  timeout 200 python scripts/bench_variants.py 2>&1 | grep -v "+pair" | tee -a gpurun_out/r2_gemm_variants_o.log
done
unset RCDM_LIB
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_o.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_o.log | cut -c1-900
