#!/bin/bash
# GEMM scheduling sweep for the prior (M = 970 rows): stream-K threshold x CTA pairing
mkdir -p gpurun_out; : > gpurun_out/prior_sweep.log
for cfg in "RCDM_SK_MIN=24 RCDM_GEMM_PAIR=1" "RCDM_SK_MIN=4 RCDM_GEMM_PAIR=1" "RCDM_SK_MIN=10 RCDM_GEMM_PAIR=1" "RCDM_SK_MIN=0 RCDM_GEMM_PAIR=1" \
           "RCDM_SK_MIN=24 RCDM_GEMM_PAIR=0" "RCDM_SK_MIN=24 RCDM_GEMM_PAIR=2" "RCDM_SK_MIN=4 RCDM_GEMM_PAIR=2" "RCDM_SK_MIN=4 RCDM_GEMM_PAIR=0"; do
  out=$(env $cfg timeout 120 python scripts/bench_prior.py --steps 50 --reps 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['achieved_tflops'],1), 'TF/s')" 2>&1)
  echo "$cfg -> $out" | tee -a gpurun_out/prior_sweep.log
done
