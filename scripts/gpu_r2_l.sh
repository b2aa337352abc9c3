#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/micro/tmem_ld_bench.out 2>&1 | tee gpurun_out/r2_tmem_ld_bench.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_ops.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -c "^ok " gpurun_out/r2_sanitizer_racecheck.log; grep -E "RACECHECK SUMMARY|FAILED|SOME|ALL OK|Error: proc" gpurun_out/r2_sanitizer_racecheck.log | tail -5
