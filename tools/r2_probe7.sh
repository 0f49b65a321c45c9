#!/bin/bash
set -x
mkdir -p gpurun_out
for v in "" _norec _nowd; do
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200$v.so timeout 200 python tools/quickcheck.py > gpurun_out/p7_quick$v.txt 2>&1
done
for v in _r1 "" _norec; do
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200$v.so timeout 120 python tools/bsweep.py classic 256,32 > gpurun_out/p7_bsweep_classic$v.txt 2>&1
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200$v.so timeout 120 python tools/bsweep.py simplified 256,32 > gpurun_out/p7_bsweep_simple$v.txt 2>&1
done
