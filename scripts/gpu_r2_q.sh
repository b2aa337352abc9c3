#!/bin/bash
# round 2, call q: epilogue with compile-time residual / statistics variants + division-free tile walk; split K / V rings in
# flash attention (V box without the padded chunk); ops parity, GEMM + flash micro-benchmarks, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=5 --timeout=300 > gpurun_out/r2_pytest_ops_q.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/r2_pytest_ops_q.log | cut -c1-300
timeout 200 python scripts/bench_variants.py 2>&1 | grep -v "+pair" | tee gpurun_out/r2_gemm_variants_q.log
timeout 200 python scripts/bench_ops.py flash 2>&1 | tee gpurun_out/r2_bench_flash_q.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/r2_pytest_gpu_q.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_q.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_q.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_q.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['derived']['ms_per_ddim_step'], d['clocks'])
print(json.dumps(d['roofline']['by_kind']))"
