#!/bin/bash
# Round-2 final validation on one B200 (after the row helpers and the K4 work list): GPU test suite, smoke, both bench arms,
# the secondary bench lines that changed, the all-configs table, batch sweeps.
cd /root/repo
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference 2>/dev/null | tail -1
} > gpurun_out/final2.txt 2>&1
python bench.py 2>/dev/null | tail -1 > gpurun_out/r2_bench_default.json
python bench.py --variant classic --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r2_bench_classic.json
python bench.py --workload cfg4 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r2_bench_cfg4.json
python bench.py --workload cfg4 --variant simplified --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r2_bench_cfg4_simplified.json
python tools/bench_configs.py > gpurun_out/r2_all_configs_device_time.txt 2>&1
python bench.py --workload cfg4full --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_bench_cfg4full.json
