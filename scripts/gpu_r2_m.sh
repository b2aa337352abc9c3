#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=5 --timeout=300 > gpurun_out/r2_pytest_ops_m.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/r2_pytest_ops_m.log | cut -c1-300
echo "--- flash: product (KV=3)"; timeout 200 python scripts/bench_ops.py flash 2>&1 | tee gpurun_out/r2_bench_flash_m.log
echo "--- flash: KV=2 variant"; RCDM_LIB=$PWD/rcdms_b200/_Cxkv2/librcdm_b200.so timeout 200 python scripts/bench_ops.py flash 2>&1 | tee -a gpurun_out/r2_bench_flash_m.log
echo "--- GEGLU_BN=160 variant"
RCDM_LIB=$PWD/rcdms_b200/_Cxgg/librcdm_b200.so timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --maxfail=5 --timeout=300 -k parity > gpurun_out/r2_pytest_ops_gg.log 2>&1
echo "pytest gg rc=$?"; tail -3 gpurun_out/r2_pytest_ops_gg.log | cut -c1-300
timeout 200 python scripts/bench_variants.py 2>&1 | grep -v "+pair" | grep geglu1
RCDM_LIB=$PWD/rcdms_b200/_Cxgg/librcdm_b200.so timeout 200 python scripts/bench_variants.py 2>&1 | grep -v "+pair" | grep geglu1
