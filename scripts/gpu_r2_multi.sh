#!/bin/bash
# Round-2 multi-GPU evidence (gpurun --gpus 2): sharded denoise == single-GPU result bitwise over NCCL; bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu_gather.py -m gpu -q --timeout=600 -rs > gpurun_out/r2_multi_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2_multi_pytest.log | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_multi_bench2.log 2>&1
echo "bench N=2 rc=$?"; tail -c 1500 gpurun_out/r2_multi_bench2.log
