#!/bin/bash
# Round-2 multi-GPU evidence on 8 GPUs (gpurun --gpus 8): sharded denoise == single-GPU result bitwise over NCCL at world sizes
# 2 / 4 / 8; bench at N = 8 (weak scaling, 1 clip per GPU; extra line: config 5, 8 clips per GPU)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m pytest tests/test_multi_gpu_gather.py -m gpu -q --timeout=900 -rs > gpurun_out/r2_multi8_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_multi8_pytest.log | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_multi8_bench.log 2>&1
echo "bench N=8 rc=$?"; tail -1 gpurun_out/r2_multi8_bench.log | cut -c1-400
