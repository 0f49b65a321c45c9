#!/bin/bash
# two GPUs: NCCL sharding test + the strong-scaling bench line (weak as secondary)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/p25_gpus.txt 2>&1
timeout 300 python -m pytest tests/test_sharding_gpu.py -m gpu -q --timeout=280 > gpurun_out/p25_pytest.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2_bench_2gpu_strong.json 2> gpurun_out/p25_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/p25_ref.json 2>> gpurun_out/p25_bench.err
