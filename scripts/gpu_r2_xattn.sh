#!/bin/bash
# round 2: short-key-range attention (cross_attn_mma_kernel) against the flash kernel - per-shape timing, op parity with both
# kernels, ncu --set full of both on 80 images x 4096 queries x 91 keys (outputs: profiles/r02_xattn_bench*.txt,
# profiles/r02_ncu_cross_attn.txt), then the A/B bench line (library option attn_short_kv = 1 | 0)
mkdir -p gpurun_out
timeout 300 python scripts/bench_xattn.py > gpurun_out/r2k_xattn_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/r2k_xattn_bench.log
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "op_parity" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2k_pytest.log)"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'cross_attn_mma' -s 2 -c 1 -f -o gpurun_out/r2o_xattn_v3 python scripts/prof_xattn.py > gpurun_out/r2o_ncu.log 2>&1; echo "ncu mma rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'flash_attn4' -s 2 -c 1 -f -o gpurun_out/r2l_xattn_flash python scripts/prof_xattn.py > gpurun_out/r2l_ncu2.log 2>&1; echo "ncu flash rc=$?"
F="--steps 3 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline --no-pixels --no-extra-configs"
timeout 300 python bench.py $F > gpurun_out/r2j_bench_short_kv.log 2>&1; echo "short rc=$?"; tail -1 gpurun_out/r2j_bench_short_kv.log | cut -c1-300
timeout 300 python bench.py $F --lib-option attn_short_kv=0 > gpurun_out/r2j_bench_flash_only.log 2>&1; echo "flash rc=$?"; tail -1 gpurun_out/r2j_bench_flash_only.log | cut -c1-300
