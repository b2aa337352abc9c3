#!/bin/bash
# round 2, call p: GroupNorm apply regridded (2 CTAs/SM); full GPU tests; bench; ncu --set full of the K = 320 GEMMs (2 producers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/r2_pytest_gpu_p.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_p.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_p.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_p.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['derived']['ms_per_ddim_step'], d['clocks'])
print(json.dumps(d['roofline']['by_kind']))"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -c 9 -f -o gpurun_out/r2_gemm_small_p python scripts/prof_gemm_small.py > gpurun_out/r2_ncu_gemm_small_p.log 2>&1
echo "ncu rc=$?"
