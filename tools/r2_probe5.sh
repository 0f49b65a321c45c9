#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/p5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p5_pytest.log
timeout 300 python tools/bsweep.py simplified 256,128,64,32 > gpurun_out/p5_bsweep_simple.txt 2>&1
timeout 300 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/p5_bsweep_classic.txt 2>&1
timeout 600 python bench.py > gpurun_out/p5_bench.json 2> gpurun_out/p5_bench.err
