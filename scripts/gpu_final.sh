#!/bin/bash
# Final evidence round: full GPU suite, smoke, per-op profile, stage-2 bench (+ reference arm), prior bench, ncu launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=180 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log | cut -c1-300
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/profile64.log 2>&1
echo "profile rc=$?"; head -14 gpurun_out/profile64.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_full.log
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; cut -c1-400 gpurun_out/bench_ref.log
timeout 400 python bench.py --workload prior --steps 3 --warmup 3 > gpurun_out/bench_prior_full.log 2>&1
echo "prior bench rc=$?"; tail -c 2500 gpurun_out/bench_prior_full.log
timeout 200 python bench.py --workload prior --impl reference > gpurun_out/bench_prior_ref.log 2>&1; cut -c1-300 gpurun_out/bench_prior_ref.log
timeout 500 ncu -k regex:'gemm_tcgen05|flash_attn|gn_fused|gn_|temporal_attn|layernorm|ddim|upsample|im2col|tokens_to|temb|gemv' \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches_dram.csv python scripts/one_forward.py 2 > gpurun_out/ncu_launches_dram.log 2>&1
python scripts/traffic_summary.py gpurun_out/launches_dram.csv 0 gpurun_out/gemm_traffic.json | tail -22
