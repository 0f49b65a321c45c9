#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/bsweep.py classic 256,200 > gpurun_out/p34_classic.txt 2>&1
CTCB200_PLAN=4,2,0,6,0 timeout 100 python tools/bsweep.py classic 256,200 >> gpurun_out/p34_classic.txt 2>&1
CTCB200_PLAN=3,2,1,6,0 timeout 100 python tools/bsweep.py classic 256 >> gpurun_out/p34_classic.txt 2>&1
CTCB200_PLAN=3,3,1,6,0 timeout 100 python tools/bsweep.py classic 256 >> gpurun_out/p34_classic.txt 2>&1
timeout 100 python tools/quickcheck.py default > gpurun_out/p34_quick.txt 2>&1
