#!/bin/bash
# round 2, last check: short-key-range attention on the mma.sync kernel by default - full GPU suite, smoke(), short bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2n_pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2n_smoke.log | cut -c1-160
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline --no-pixels --no-extra-configs > gpurun_out/r2n_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2n_bench.log | cut -c1-260
