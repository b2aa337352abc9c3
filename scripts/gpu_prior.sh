#!/bin/bash
# GPU round trip for the stage-1 prior: parity tests (+ the stage-2 suite, whose GEMM epilogue gained an activation),
# prior bench, launch list.  Every command under `timeout`.  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prior_gpu.py -m gpu -q --timeout=180 > gpurun_out/pytest_prior.log 2>&1
echo "pytest prior rc=$?"; tail -40 gpurun_out/pytest_prior.log | cut -c1-400
RCDM_PRIOR_SIMPLE=1 timeout 400 python -m pytest tests/test_prior_gpu.py -m gpu -q --timeout=180 -k "forward_matches or sampling_loop" > gpurun_out/pytest_prior_simple.log 2>&1
echo "pytest prior (simple GEMM) rc=$?"; tail -8 gpurun_out/pytest_prior_simple.log | cut -c1-400
timeout 420 python -m pytest tests -m gpu -q -x --timeout=120 --ignore=tests/test_prior_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest stage-2 rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 300 python scripts/bench_prior.py --steps 100 --reps 3 > gpurun_out/bench_prior.log 2>&1
echo "bench prior rc=$?"; tail -3 gpurun_out/bench_prior.log | cut -c1-1500
timeout 200 python scripts/bench_prior.py --steps 100 --reps 2 --no-graph > gpurun_out/bench_prior_nograph.log 2>&1
tail -1 gpurun_out/bench_prior_nograph.log | cut -c1-600
timeout 400 ncu -k regex:'gemm_tcgen05|masked_attn|layernorm_kernel|temporal_attn|prior_assemble|unclip_cfg' \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches_prior.csv python scripts/bench_prior.py --once --layers 4 > gpurun_out/ncu_prior.log 2>&1
python scripts/traffic_summary.py gpurun_out/launches_prior.csv 0 gpurun_out/prior_traffic.json 2>&1 | tail -25
