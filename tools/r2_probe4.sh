#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/p4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p4_pytest.log
timeout 300 python tools/bsweep.py simplified 256,128,64,32 > gpurun_out/p4_bsweep_simple.txt 2>&1
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_atomics.so timeout 300 python tools/bsweep.py simplified 256,128,64,32 > gpurun_out/p4_bsweep_simple_atomics.txt 2>&1
timeout 300 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/p4_bsweep_classic.txt 2>&1
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_atomics.so timeout 300 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/p4_bsweep_classic_atomics.txt 2>&1
CTCB200_FUSED_W=8 CTCB200_TIMING_B=32 CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_timing.so timeout 300 python tools/fused_timing.py > gpurun_out/p4_timing_simple_B32.txt 2>&1
CTCB200_TIMING_B=256 CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_timing.so timeout 300 python tools/fused_timing.py > gpurun_out/p4_timing_simple_B256.txt 2>&1
