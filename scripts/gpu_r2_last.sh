#!/bin/bash
# last validation of the round: sanitizers over the fused feed-forward kernel, smoke(), default bench line
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 10 python scripts/sanitize_ops.py ffn > gpurun_out/r2_sanitizer_ffn_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "^ok |FAILED|ALL OK|SOME|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_ffn_$tool.log | tail -4
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_last_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r2_last_smoke.log | cut -c1-200
timeout 1200 python bench.py > gpurun_out/r2_last_bench.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_last_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['derived']['ms_per_ddim_step'], d['clocks'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'])"
