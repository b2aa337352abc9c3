#!/bin/bash
# round 2, call y: flash attention with loader warp + register-resident descriptors; polynomial-exp share 2/8, 3/8, 4/8
mkdir -p gpurun_out
: > gpurun_out/r2_bench_flash_y.log
for v in "" _Cxp2 _Cxp4; do
  if [ -n "$v" ]; then export RCDM_LIB=$PWD/rcdms_b200/$v/librcdm_b200.so; fi
  echo "--- ${v:-product (3/8)}" | tee -a gpurun_out/r2_bench_flash_y.log
  timeout 200 python scripts/bench_ops.py flash 2>&1 | tee -a gpurun_out/r2_bench_flash_y.log
done
unset RCDM_LIB
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/r2_pytest_gpu_y.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_y.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_y.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_y.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['derived']['ms_per_ddim_step'], d['clocks'])
print(json.dumps(d['roofline']['by_kind']))"
