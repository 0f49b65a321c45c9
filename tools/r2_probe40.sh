#!/bin/bash
cd /root/repo
export CTCB200_TIMING_B=256 CTCB200_TIMING_TVL=1600,5000,400 CTCB200_FUSED_W=1
{
echo "== HEAD"; CTCB200_LIB=/root/repo/tf_seq2seq_losses_b200/libctc_b200_timing.so python tools/fused_timing.py classic
echo "== r1"; (cd .r1tmp && CTCB200_LIB=/root/repo/.r1tmp/tf_seq2seq_losses_b200/libctc_b200_timing.so python tools/fused_timing.py classic)
} > gpurun_out/p40.txt 2>&1
