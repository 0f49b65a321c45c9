#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in "" _old; do
  echo "== lib '$v' rep $rep" >> gpurun_out/p35_ab.txt
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200$v.so timeout 100 python tools/bsweep.py classic 256 2>&1 | grep "B=" >> gpurun_out/p35_ab.txt
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200$v.so timeout 100 python tools/bsweep.py simplified 256 2>&1 | grep "B=" >> gpurun_out/p35_ab.txt
done
done
