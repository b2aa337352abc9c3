#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_prior_gpu.py -m gpu -q --timeout=180 > gpurun_out/pytest_prior.log 2>&1
echo "pytest prior rc=$?"; tail -4 gpurun_out/pytest_prior.log | cut -c1-300
timeout 300 python bench.py --workload prior --steps 3 --warmup 3 > gpurun_out/bench_prior_full.log 2>&1
echo "prior rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_prior_full.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['cpu_baseline'], d['gpu_launches'])"
timeout 300 ncu --set full --clock-control none -k regex:'masked_attn_mma|temporal_attn_wide|layernorm_kernel' -c 9 \
  -o gpurun_out/prior_small_kernels -f python scripts/bench_prior.py --once --layers 2 > gpurun_out/ncu_prior_small.log 2>&1
echo "ncu small rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:'gemm_tcgen05' --launch-skip 8 -c 13 \
  -o gpurun_out/prior_gemm -f python scripts/bench_prior.py --once --layers 2 > gpurun_out/ncu_prior_gemm.log 2>&1
echo "ncu gemm rc=$?"; ls -la gpurun_out/*.ncu-rep
