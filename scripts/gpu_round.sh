#!/bin/bash
# One GPU round trip: parity tests, per-op forward profile, op micro-benches, bench line.  Outputs -> gpurun_out/
# Every command runs under `timeout` so that a hung kernel costs minutes, not the whole call.
# usage: bash scripts/gpu_round.sh [ncu] [VAR=val ...]   ("ncu" adds the ncu launch list of one forward)
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --timeout=120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-600
RCDM_SK_MIN=1 timeout 420 python -m pytest tests -m gpu -q -x --timeout=120 > gpurun_out/pytest_gpu_sk1.log 2>&1
echo "pytest(RCDM_SK_MIN=1) rc=$?"; tail -4 gpurun_out/pytest_gpu_sk1.log | cut -c1-600
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/profile64.log 2>&1
echo "profile rc=$?"; head -12 gpurun_out/profile64.log
timeout 200 python scripts/bench_gemm_shapes.py > gpurun_out/bench_gemm_shapes.log 2>&1
cat gpurun_out/bench_gemm_shapes.log
timeout 200 python scripts/bench_ops.py > gpurun_out/bench_ops.log 2>&1
grep -i "norm\|temporal\|copy" gpurun_out/bench_ops.log
timeout 200 python scripts/bench_ops.py flash 2>&1 | tee gpurun_out/bench_flash.log
RCDM_ATTN_V=3 timeout 200 python scripts/bench_ops.py flash 2>&1 | tee gpurun_out/bench_flash_v3.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.log
for extra in "$@"; do
  if [ "$extra" = "ncudram" ]; then
    timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches_dram.csv python scripts/one_forward.py 2 > gpurun_out/ncu_launches_dram.log 2>&1
    python scripts/traffic_summary.py gpurun_out/launches_dram.csv 0 gpurun_out/gemm_traffic.json | tail -30
    continue
  fi
  if [ "$extra" = "fullbench" ]; then
    timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1; tail -c 1200 gpurun_out/bench_full.log
    timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; cat gpurun_out/bench_ref.log | cut -c1-600
    continue
  fi
  if [ "$extra" = "ncu" ]; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python scripts/one_forward.py 2 > gpurun_out/ncu_launches.log 2>&1
    python scripts/launch_summary.py gpurun_out/launches.csv 800 | head -30
    continue
  fi
  echo "== $extra"
  env $extra timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/_v.log 2>&1
  echo "$extra $(cat gpurun_out/_v.log)" >> gpurun_out/bench_variants.log
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/_v.log").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"], 2), "frames/s", round(d["value"], 3), d["clocks"])
except Exception as e:
    print("variant failed:", e, open("gpurun_out/_v.log").read()[-400:])
PY
done
