#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=180 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print({k:(round(v['ms'],3), round(v['tflops'],1), round(v['gbs'],0)) for k,v in d['roofline']['by_kind'].items()}); print(d['cpu_baseline'])"
