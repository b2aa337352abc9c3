#!/bin/bash
# Round-2 FINAL evidence run of the last session (LayerNorm / proj_out folds in): outputs -> gpurun_out/r2z_*, copied to profiles/r02_* afterwards
# ncu launch list with DRAM bytes of one forward -> traffic json (read by bench.py), full GPU test suite, smoke(), default
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2z_launches_dram.csv python scripts/one_forward.py 2 > gpurun_out/r2z_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/traffic_summary.py gpurun_out/r2z_launches_dram.csv 0 gpurun_out/r2z_gemm_traffic.json > gpurun_out/r2z_launches_dram.txt 2>&1  # (whole process; the forward-only summary in profiles/ was cut from the csv afterwards: skip = first launch of the second forward)
cp gpurun_out/r2z_gemm_traffic.json profiles/r02_gemm_traffic.json; tail -22 gpurun_out/r2z_launches_dram.txt | cut -c1-170
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2z_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest_gpu.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
echo "smoke rc=$?"; tail -4 gpurun_out/r2z_smoke.log | cut -c1-200
timeout 1200 python bench.py > gpurun_out/r2z_bench.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2z_bench.log | cut -c1-400
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.log 2>&1
echo "bench ref rc=$?"; tail -c 400 gpurun_out/r2z_bench_ref.log
timeout 240 python scripts/profile_forward.py 64 1 > gpurun_out/r2z_profile64.log 2>&1
echo "profile rc=$?"; head -3 gpurun_out/r2z_profile64.log
timeout 900 python bench.py --workload prior > gpurun_out/r2z_prior_bench.log 2>&1
echo "prior bench rc=$?"; tail -1 gpurun_out/r2z_prior_bench.log | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -c 12 -f -o gpurun_out/r2z_gemm_k320 python scripts/prof_gemm_small.py > gpurun_out/r2z_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
