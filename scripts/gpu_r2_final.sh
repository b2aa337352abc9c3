#!/bin/bash
# Round-2 evidence run on one B200: full GPU test suite, smoke(), default bench line (both arms), ncu launch list with DRAM bytes
# of one forward, ncu --set full of the streaming GroupNorm / a deep-K conv.  Outputs -> gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2_final_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest_gpu.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
echo "smoke rc=$?"; tail -6 gpurun_out/r2_final_smoke.log | cut -c1-200
timeout 1200 python bench.py > gpurun_out/r2_final_bench.log 2>&1
echo "bench rc=$?"; tail -c 600 gpurun_out/r2_final_bench.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.log 2>&1
echo "bench ref rc=$?"; tail -c 700 gpurun_out/r2_final_bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches_dram.csv python scripts/one_forward.py 2 > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/traffic_summary.py gpurun_out/r2_launches_dram.csv 0 gpurun_out/r2_gemm_traffic.json > gpurun_out/r2_launches_dram.txt 2>&1; tail -25 gpurun_out/r2_launches_dram.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gn_apply_stats -s 20 -c 3 -o gpurun_out/r2_gn_apply python scripts/one_forward.py 1 > gpurun_out/r2_ncu_gn.log 2>&1
echo "ncu gn rc=$?"
