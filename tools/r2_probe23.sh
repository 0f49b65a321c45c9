#!/bin/bash
# round-2 profile run: bench lines for the secondary workloads, the all-configs table, ncu launch list and full captures
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline"
timeout 300 $B --variant classic --no-e2e > gpurun_out/r2_bench_classic.json 2> gpurun_out/p23.err
timeout 300 $B --ragged --no-e2e > gpurun_out/r2_bench_ragged.json 2>> gpurun_out/p23.err
timeout 300 $B --workload cfg1 --no-e2e > gpurun_out/r2_bench_cfg1.json 2>> gpurun_out/p23.err
timeout 300 $B --workload cfg4 --no-e2e > gpurun_out/r2_bench_cfg4.json 2>> gpurun_out/p23.err
timeout 300 $B --dtype-in bf16 > gpurun_out/r2_bench_bf16in.json 2>> gpurun_out/p23.err
timeout 300 $B --dtype-in bf16 --dtype-grad bf16 > gpurun_out/r2_bench_bf16.json 2>> gpurun_out/p23.err
timeout 400 python tools/bench_configs.py > gpurun_out/r2_all_configs_device_time.txt 2>> gpurun_out/p23.err
timeout 200 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/r2_bsweep_classic.txt 2>> gpurun_out/p23.err
timeout 200 python tools/bsweep.py simplified 256,128,64,32,16 > gpurun_out/r2_bsweep_simplified.txt 2>> gpurun_out/p23.err
# ncu: launch list of the bench command, then one full capture per kernel of interest (never a bench number)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2_simplified.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>> gpurun_out/p23.err
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:kf_fused -s 3 -c 1 -o gpurun_out/r2_fused_simple_B256 python tools/bsweep.py simplified 256 > /dev/null 2>> gpurun_out/p23.err
$NCU -k regex:kf_fused -s 3 -c 1 -o gpurun_out/r2_fused_classic_B256 python tools/bsweep.py classic 256 > /dev/null 2>> gpurun_out/p23.err
$NCU -k regex:kf_fused -s 3 -c 1 -o gpurun_out/r2_fused_simple_B32 python tools/bsweep.py simplified 32 > /dev/null 2>> gpurun_out/p23.err
$NCU -k regex:kf_fused -s 3 -c 1 -o gpurun_out/r2_fused_classic_B32 python tools/bsweep.py classic 32 > /dev/null 2>> gpurun_out/p23.err
ls -la gpurun_out | tail -20
