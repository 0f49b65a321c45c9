#!/bin/bash
# round-2 probe 1: sanity of the rebuilt library + where the fused kernel's warps wait at B=32 / B=256
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p1_pytest.log
for B in 256 32; do
  CTCB200_TIMING_B=$B CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_timing.so timeout 300 python tools/fused_timing.py > gpurun_out/p1_timing_simple_B$B.txt 2>&1
  CTCB200_TIMING_B=$B CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_timing.so timeout 300 python tools/fused_timing.py classic > gpurun_out/p1_timing_classic_B$B.txt 2>&1
done
timeout 300 python tools/bsweep.py > gpurun_out/p1_bsweep.txt 2>&1
