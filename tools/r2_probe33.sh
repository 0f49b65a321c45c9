#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --workload cfg4full --steps 5 --warmup 3 --no-secondary > gpurun_out/r2_bench_cfg4full_4gpu_strong.json 2> gpurun_out/p33_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload cfg4full --steps 5 --warmup 3 --no-secondary > gpurun_out/r2_bench_cfg4full_2gpu_strong.json 2>> gpurun_out/p33_bench.err
