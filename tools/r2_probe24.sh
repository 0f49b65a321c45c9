#!/bin/bash
mkdir -p gpurun_out
CTCB200_TVL=1600,5000,400 timeout 200 python tools/bsweep.py classic 256 > gpurun_out/p24_cfg4_default.txt 2>&1
CTCB200_TVL=1600,5000,400 CTCB200_PLAN=2,2,0,4,0 timeout 200 python tools/bsweep.py classic 256 > gpurun_out/p24_cfg4_w2.txt 2>&1
CTCB200_TVL=1600,5000,400 CTCB200_PLAN=1,2,0,2,0 timeout 200 python tools/bsweep.py classic 256 > gpurun_out/p24_cfg4_w1.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-e2e --dtype-in bf16 > gpurun_out/p24_bench_bf16in.json 2> gpurun_out/p24.err
timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -q --timeout=300 -k "bfloat16 or graph_capture or readme or README" > gpurun_out/p24_pytest.log 2>&1
timeout 200 python tools/bench_configs.py 2>&1 | head -3 > gpurun_out/p24_readme.txt
