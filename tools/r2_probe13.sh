#!/bin/bash
mkdir -p gpurun_out
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_wd.so timeout 100 python tools/dbg_case.py 8,64,10,30 1 2>&1 | grep -v "^  " | tail -150 > gpurun_out/p13_a.txt
