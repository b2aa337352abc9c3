#!/bin/bash
# round 2: short-key-range attention on the register-resident mma.sync kernel - parity + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2j_pytest.log)"
grep -i "short1\|short0" gpurun_out/r2j_pytest.log | head -5
F="--steps 3 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline --no-pixels --no-extra-configs"
timeout 300 python bench.py $F > gpurun_out/r2j_bench_short_kv.log 2>&1; echo "short rc=$?"; tail -1 gpurun_out/r2j_bench_short_kv.log | cut -c1-300
timeout 300 python bench.py $F --lib-option attn_short_kv=0 > gpurun_out/r2j_bench_flash_only.log 2>&1; echo "flash rc=$?"; tail -1 gpurun_out/r2j_bench_flash_only.log | cut -c1-300
