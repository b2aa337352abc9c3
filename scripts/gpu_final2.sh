#!/bin/bash
# Final evidence round (round 1): full GPU suite, smoke, stage-2 bench (+ A/B of the wide temporal kernel for the UNet
# head dims), prior bench, reference arms, ncu launch lists.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=180 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
RCDM_TEMPORAL_WIDE_ALL=1 timeout 400 python -m pytest tests -m gpu -q --timeout=180 -k "temporal or unet or pipeline or golden" > gpurun_out/pytest_gpu_wide_all.log 2>&1
echo "pytest (RCDM_TEMPORAL_WIDE_ALL=1) rc=$?"; tail -3 gpurun_out/pytest_gpu_wide_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print({k:(round(v['ms'],3), round(v['tflops'],1), round(v['gbs'],0)) for k,v in d['roofline']['by_kind'].items()}); print(d['cpu_baseline'])"
RCDM_TEMPORAL_WIDE_ALL=1 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wide_all.log 2>&1
echo "bench (RCDM_TEMPORAL_WIDE_ALL=1) rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_wide_all.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']); print({k:(round(v['ms'],3), round(v['gbs'],0)) for k,v in d['roofline']['by_kind'].items() if 'temporal' in k})"
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; cut -c1-200 gpurun_out/bench_ref.log
timeout 400 python bench.py --workload prior --steps 3 --warmup 3 > gpurun_out/bench_prior_full.log 2>&1
echo "prior rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_prior_full.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['cpu_baseline'], d['gpu_launches'])"
timeout 500 ncu -k regex:'gemm_tcgen05|flash_attn|gn_fused|gn_|temporal_attn|layernorm|ddim|upsample|im2col|tokens_to|temb|gemv' \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches_dram.csv python scripts/one_forward.py 2 > gpurun_out/ncu_launches_dram.log 2>&1
python scripts/traffic_summary.py gpurun_out/launches_dram.csv 0 gpurun_out/gemm_traffic.json | tail -20 | cut -c1-160
timeout 300 ncu -k regex:'gemm_tcgen05|masked_attn|layernorm|temporal_attn|prior_assemble|unclip_cfg' \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches_prior.csv python scripts/bench_prior.py --once --layers 4 > gpurun_out/ncu_prior.log 2>&1
python scripts/traffic_summary.py gpurun_out/launches_prior.csv 0 gpurun_out/prior_traffic.json | head -9 | cut -c1-160
