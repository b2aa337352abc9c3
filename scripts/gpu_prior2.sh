#!/bin/bash
# second prior round trip: parity of the new kernels, bench, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prior_gpu.py -m gpu -q --timeout=180 > gpurun_out/pytest_prior.log 2>&1
echo "pytest prior rc=$?"; tail -40 gpurun_out/pytest_prior.log | cut -c1-400
timeout 100 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout=120 > gpurun_out/pytest_ops.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/pytest_ops.log | cut -c1-300
timeout 300 python scripts/bench_prior.py --steps 100 --reps 3 > gpurun_out/bench_prior.log 2>&1
echo "bench prior rc=$?"; tail -1 gpurun_out/bench_prior.log | cut -c1-700
RCDM_MASKED_ATTN_MMA=0 timeout 200 python scripts/bench_prior.py --steps 100 --reps 2 > gpurun_out/bench_prior_nomma.log 2>&1
echo "mma off:"; tail -1 gpurun_out/bench_prior_nomma.log | cut -c1-400
RCDM_TEMPORAL_WIDE=0 timeout 200 python scripts/bench_prior.py --steps 100 --reps 2 > gpurun_out/bench_prior_nowide.log 2>&1
echo "wide temporal off:"; tail -1 gpurun_out/bench_prior_nowide.log | cut -c1-400
timeout 300 ncu -k regex:'gemm_tcgen05|masked_attn|layernorm_kernel|temporal_attn|prior_assemble|unclip_cfg' \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches_prior.csv python scripts/bench_prior.py --once --layers 4 > gpurun_out/ncu_prior.log 2>&1
python scripts/traffic_summary.py gpurun_out/launches_prior.csv 0 gpurun_out/prior_traffic.json 2>&1 | tail -12
