#!/bin/bash
# flakiness check: the full GPU suite three times in a row + smoke()
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/r2_repeat_$i.log 2>&1
  echo "run $i rc=$? $(tail -1 gpurun_out/r2_repeat_$i.log | cut -c1-120)"
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_repeat_smoke.log 2>&1
echo "smoke rc=$?"
