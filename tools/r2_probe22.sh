#!/bin/bash
mkdir -p gpurun_out
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_wdog.so timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 --timeout=300 > gpurun_out/p22_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p22_pytest.log
timeout 200 python tools/bsweep.py simplified 256,128,64,32 > gpurun_out/p22_bsweep_simple.txt 2>&1
CTCB200_PLAN=8,3,1,16,5 timeout 200 python tools/bsweep.py simplified 64,32 > gpurun_out/p22_bsweep_simple_noidle.txt 2>&1
timeout 200 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/p22_bsweep_classic.txt 2>&1
CTCB200_PLAN=8,3,1,16,5 timeout 200 python tools/bsweep.py classic 64,32 > gpurun_out/p22_bsweep_classic_noidle.txt 2>&1
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_l2hints.so timeout 200 python tools/bsweep.py simplified 256,128 > gpurun_out/p22_bsweep_simple_l2hints.txt 2>&1
CTCB200_TVL=500,29,100 timeout 100 python tools/bsweep.py classic 32 > gpurun_out/p22_cfg1.txt 2>&1
CTCB200_TVL=500,29,100 CTCB200_FLAGS=4 timeout 100 python tools/bsweep.py classic 32 >> gpurun_out/p22_cfg1.txt 2>&1
timeout 600 python bench.py > gpurun_out/p22_bench.json 2> gpurun_out/p22_bench.err
