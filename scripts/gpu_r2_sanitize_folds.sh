#!/bin/bash
# round 2, last session: compute-sanitizer over the new kernels / launch forms (LayerNorm-folded GEMM chains with 32 statistic
# parts, fold_proj_kernel + two-segment GEMM, cross_attn_mma_kernel, tiny UNet + tiny prior forwards through them)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_ops.py folds > gpurun_out/r2m_sanitizer_folds_$tool.log 2>&1
  echo "$tool rc=$?"; grep -c "^ok " gpurun_out/r2m_sanitizer_folds_$tool.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|FAILED|SOME|ALL OK|Error: proc" gpurun_out/r2m_sanitizer_folds_$tool.log | tail -5
done
