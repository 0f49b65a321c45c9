#!/bin/bash
cd /root/repo
D=tf_seq2seq_losses_b200
{
python tools/ab_lib.py classic 256,1600,5000,400 $D/libctc_b200.so $D/libctc_b200_r1.so
python tools/ab_lib.py simplified 256,1600,5000,400 $D/libctc_b200.so $D/libctc_b200_r1.so
} > gpurun_out/p39.txt 2>&1
