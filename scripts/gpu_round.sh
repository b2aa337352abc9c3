#!/bin/bash
# One GPU round trip: parity tests, per-op forward profile, op micro-benches, bench line.  Outputs -> gpurun_out/
# Every command runs under `timeout` so that a hung kernel costs minutes, not the whole call.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --timeout=120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-600
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/profile64.log 2>&1
echo "profile rc=$?"; head -12 gpurun_out/profile64.log
timeout 200 python scripts/bench_ops.py > gpurun_out/bench_ops.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.log
for extra in "$@"; do
  echo "== $extra"
  env $extra timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile 2>&1 | cut -c1-300 | tee -a gpurun_out/bench_variants.log
done
