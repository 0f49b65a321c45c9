#!/bin/bash
mkdir -p gpurun_out
CTCB200_TVL=500,29,100 timeout 100 python tools/bsweep.py classic 32 > gpurun_out/p28_cfg1.txt 2>&1
CTCB200_TVL=500,29,100 CTCB200_FLAGS=4 timeout 100 python tools/bsweep.py classic 32,64,128 >> gpurun_out/p28_cfg1.txt 2>&1
CTCB200_TVL=500,29,100 CTCB200_FLAGS=2 timeout 100 python tools/bsweep.py classic 64,128 >> gpurun_out/p28_cfg1.txt 2>&1
CTCB200_TVL=255,32,127 CTCB200_FLAGS=4 timeout 100 python tools/bsweep.py classic 256,32 >> gpurun_out/p28_cfg1.txt 2>&1
CTCB200_TVL=255,32,127 CTCB200_FLAGS=2 timeout 100 python tools/bsweep.py classic 256,32 >> gpurun_out/p28_cfg1.txt 2>&1
timeout 300 python tools/quickcheck.py default > gpurun_out/p28_quick.txt 2>&1
timeout 300 python tools/quickcheck.py "W2 R4" >> gpurun_out/p28_quick.txt 2>&1
