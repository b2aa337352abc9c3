#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/micro/tma_box_bench.out 2>&1 | head -4 | tee gpurun_out/r2_tma_interference.log
