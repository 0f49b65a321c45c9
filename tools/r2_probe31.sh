#!/bin/bash
mkdir -p gpurun_out
CTCB200_TVL=1600,5000,400 timeout 200 python tools/bsweep.py classic 256 > gpurun_out/p31_cfg4.txt 2>&1
CTCB200_TVL=1600,5000,400 CTCB200_PLAN=1,2,0,2,8 timeout 200 python tools/bsweep.py classic 256 >> gpurun_out/p31_cfg4.txt 2>&1
CTCB200_TVL=1600,5000,400 CTCB200_PLAN=2,2,0,4,0 timeout 200 python tools/bsweep.py classic 256 >> gpurun_out/p31_cfg4.txt 2>&1
for plan in "W2 R4" "W1 R1" "default"; do timeout 200 python tools/quickcheck.py "$plan" >> gpurun_out/p31_quick.txt 2>&1; done
