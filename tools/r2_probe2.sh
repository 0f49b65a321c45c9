#!/bin/bash
# round-2 probe 2: ncu --set full captures (with source) of the fused kernel at B=32, K2 at configs[1], K4 at configs[3]
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:kf_fused -s 3 -c 1 -o gpurun_out/r2a_fused_simple_B32 python tools/bsweep.py simplified 32 > gpurun_out/p2_a.log 2>&1
$NCU -k regex:kf_fused -s 3 -c 1 -o gpurun_out/r2a_fused_classic_B32 python tools/bsweep.py classic 32 > gpurun_out/p2_b.log 2>&1
$NCU -k regex:k2_recursion -s 2 -c 1 -o gpurun_out/r2a_k2_cfg1 python bench.py --workload cfg1 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/p2_c.log 2>&1
$NCU -k regex:k4_hessian -s 2 -c 1 -o gpurun_out/r2a_k4_cfg3 python -c "import sys; sys.path.insert(0,'tools'); import bench_configs as b; b.hessian_case(64,50,32,15)" > gpurun_out/p2_d.log 2>&1
ls -la gpurun_out
