#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_gemm_variants_j.log
timeout 200 python scripts/bench_variants.py >> gpurun_out/r2_gemm_variants_j.log 2>&1
RCDM_LIB=$PWD/rcdms_b200/_Cxg4/librcdm_b200.so timeout 200 python scripts/bench_variants.py >> gpurun_out/r2_gemm_variants_j.log 2>&1
grep -v "+pair" gpurun_out/r2_gemm_variants_j.log
RCDM_LIB=$PWD/rcdms_b200/_Cxg4/librcdm_b200.so timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q --maxfail=5 --timeout=300 > gpurun_out/r2_pytest_g4.log 2>&1
echo "pytest g4 rc=$?"; tail -5 gpurun_out/r2_pytest_g4.log | cut -c1-300
RCDM_LIB=$PWD/rcdms_b200/_Cxg4/librcdm_b200.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs --no-eager-gpu-baseline > gpurun_out/r2_bench_g4.log 2>&1
echo "bench g4 rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_g4.log') if l.startswith('{')][-1])
print(d['value'], d['derived']['ms_per_ddim_step'], d['clocks'])
for k,v in d['roofline']['by_kind'].items(): print(k,v)
PY
