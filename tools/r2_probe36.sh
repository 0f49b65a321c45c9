#!/bin/bash
# wide label-state shapes (U up to 1024) on the staged kernels: parity tests + device time
cd /root/repo
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "size_limits or no_cpu_fallback" 2>&1 | tail -4
CTCB200_TVL=1200,1024,1023 python tools/bsweep.py classic 64,256 2>&1 | grep "B="
CTCB200_TVL=1200,1024,1023 python tools/bsweep.py simplified 64,256 2>&1 | grep "B="
CTCB200_TVL=1200,1024,511 python tools/bsweep.py classic 64,256 2>&1 | grep "B="
CTCB200_TVL=1200,1024,511 CTCB200_FLAGS=2 python tools/bsweep.py classic 64,256 2>&1 | grep "B="
} > gpurun_out/p36.txt 2>&1
