#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_prior_gpu.py tests/test_ops_gpu.py -m gpu -q --timeout=180 > gpurun_out/pytest_verify.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_verify.log | cut -c1-300
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print({k:(round(v['ms'],3), round(v['tflops'],1)) for k,v in d['roofline']['by_kind'].items()}); print(d['cpu_baseline'])"
timeout 300 python bench.py --workload prior --steps 3 --warmup 3 > gpurun_out/bench_prior_full.log 2>&1
echo "prior rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_prior_full.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['cpu_baseline'], d['gpu_launches'])"
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/profile64.log 2>&1; head -8 gpurun_out/profile64.log
