#!/bin/bash
# Evidence round: full GPU suite (+ the tiled temporal kernel), smoke, stage-2 bench, eager-GPU baselines, prior bench,
# reference arm, ncu launch lists.  ~5 minutes on one B200.  Outputs -> gpurun_out/ (copy what matters to profiles/).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=180 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
RCDM_TEMPORAL_WIDE_ALL=0 timeout 400 python -m pytest tests -m gpu -q --timeout=180 -k "temporal or unet or pipeline or golden" > gpurun_out/pytest_gpu_tiled_temporal.log 2>&1
echo "pytest (RCDM_TEMPORAL_WIDE_ALL=0: tiled temporal kernel) rc=$?"; tail -3 gpurun_out/pytest_gpu_tiled_temporal.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print({k:(round(v['ms'],3), round(v['tflops'],1), round(v['gbs'],0)) for k,v in d['roofline']['by_kind'].items()}); print(d['cpu_baseline'])"
timeout 100 python bench.py --eager-gpu-only --steps 3 > gpurun_out/bench_eager_gpu.log 2>&1; cut -c1-300 gpurun_out/bench_eager_gpu.log
timeout 100 python bench.py --eager-gpu-only --workload prior --steps 3 > gpurun_out/bench_prior_eager_gpu.log 2>&1; cut -c1-300 gpurun_out/bench_prior_eager_gpu.log
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
