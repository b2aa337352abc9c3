#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_xattn.py > gpurun_out/r2k_xattn_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/r2k_xattn_bench.log
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "op_parity" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2k_pytest.log)"
