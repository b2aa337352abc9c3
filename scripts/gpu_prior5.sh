#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_prior_gpu.py -m gpu -q --timeout=180 > gpurun_out/pytest_prior.log 2>&1
echo "pytest prior rc=$?"; tail -4 gpurun_out/pytest_prior.log | cut -c1-300
for v in 1 0; do
RCDM_LN_WIDE=$v timeout 300 python bench.py --workload prior --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prior_ln$v.log 2>&1
echo "LN_WIDE=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_prior_ln$v.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['gpu_launches'])"
done
