#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/bsweep.py simplified 148,128,96,75,74 > gpurun_out/p27_simple.txt 2>&1
timeout 200 python tools/bsweep.py classic 148,128,96,75 > gpurun_out/p27_classic.txt 2>&1
CTCB200_PLAN=7,2,1,14,1 timeout 100 python tools/bsweep.py simplified 128 > gpurun_out/p27_w7.txt 2>&1
CTCB200_PLAN=4,2,1,8,0 timeout 100 python tools/bsweep.py classic 128 > gpurun_out/p27_classic_nonsplit.txt 2>&1
timeout 400 python tools/quickcheck.py default > gpurun_out/p27_quick.txt 2>&1
