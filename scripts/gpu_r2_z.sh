#!/bin/bash
# round 2, call z: 192-wide GEMM tiles where they save a wave
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=5 --timeout=300 > gpurun_out/r2_pytest_ops_z.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/r2_pytest_ops_z.log | cut -c1-300
timeout 200 python scripts/bench_variants.py 2>&1 | grep -v "+pair" | tee gpurun_out/r2_gemm_variants_z.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/r2_pytest_gpu_z.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_z.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_z.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_z.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['derived']['ms_per_ddim_step'], d['clocks'])
print(json.dumps(d['roofline']['by_kind']))"
