#!/bin/bash
# round-2 probe 3: atomics-free scatter + split (two-CTA cluster) plan: parity, batch sweep, wait-time breakdown
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/p3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p3_pytest.log
timeout 300 python tools/bsweep.py simplified 256,192,128,64,32,16 > gpurun_out/p3_bsweep_simple.txt 2>&1
timeout 300 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/p3_bsweep_classic.txt 2>&1
for B in 32; do
  CTCB200_FUSED_W=8 CTCB200_TIMING_B=$B CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_timing.so timeout 300 python tools/fused_timing.py > gpurun_out/p3_timing_simple_B$B.txt 2>&1
  CTCB200_FUSED_W=8 CTCB200_TIMING_B=$B CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_timing.so timeout 300 python tools/fused_timing.py classic > gpurun_out/p3_timing_classic_B$B.txt 2>&1
done
CTCB200_TIMING_B=256 CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_timing.so timeout 300 python tools/fused_timing.py > gpurun_out/p3_timing_simple_B256.txt 2>&1
