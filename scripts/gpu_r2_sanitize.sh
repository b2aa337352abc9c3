#!/bin/bash
# round 2: compute-sanitizer over every kernel family of the FINAL code (two-producer GEMM, 192-wide tiles, loader-warp flash
# attention with separate K / V rings, regridded GroupNorm apply)
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_ops.py > gpurun_out/r2_sanitizer2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -c "^ok " gpurun_out/r2_sanitizer2_$tool.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|FAILED|SOME|ALL OK|Error: proc" gpurun_out/r2_sanitizer2_$tool.log | tail -5
done
