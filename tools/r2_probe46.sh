#!/bin/bash
# does the helper code cost the narrow-row kernels anything?  HEAD against the library of the commit before the helpers
cd /root/repo
D=tf_seq2seq_losses_b200
{
python tools/ab_lib.py classic 256,1000,1024,200 $D/libctc_b200.so $D/libctc_b200_pre.so
python tools/ab_lib.py simplified 256,1000,1024,200 $D/libctc_b200.so $D/libctc_b200_pre.so
python tools/ab_lib.py classic 32,1000,1024,200 $D/libctc_b200.so $D/libctc_b200_pre.so
} > gpurun_out/p46.txt 2>&1
