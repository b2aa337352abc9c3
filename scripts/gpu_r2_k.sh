#!/bin/bash
# compute-sanitizer at tiny shapes: memcheck, racecheck, synccheck over every kernel family (scripts/sanitize_ops.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --timeout=300 > gpurun_out/r2_pytest_gpu_k.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_k.log | cut -c1-300
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_ops.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -c "^ok " gpurun_out/r2_sanitizer_$tool.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|FAILED|SOME|ALL OK" gpurun_out/r2_sanitizer_$tool.log | tail -5
done
