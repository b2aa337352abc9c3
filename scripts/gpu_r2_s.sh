#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/micro/store_bench.out 2>&1 | tee gpurun_out/r2_store_bench.log
