#!/bin/bash
# pipelined phase-B scatter (load pass + compare-and-swap pass) against the serial shared-memory atomics and the round-1 library
cd /root/repo
D=tf_seq2seq_losses_b200
{
timeout 300 python tools/quickcheck.py default 2>&1 | tail -20
for s in 256,1600,5000,400 256,1000,1024,200 32,1000,1024,200; do
python tools/ab_lib.py classic $s $D/libctc_b200.so $D/libctc_b200_noscat.so $D/libctc_b200_r1.so
python tools/ab_lib.py simplified $s $D/libctc_b200.so $D/libctc_b200_noscat.so $D/libctc_b200_r1.so
done
} > gpurun_out/p41.txt 2>&1
