#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/dbg_case.py 90,6,8,3 0 2>&1 | head -30 > gpurun_out/p11_a.txt
timeout 60 python tools/dbg_case.py 3,12,5,3 0 2>&1 | head -30 > gpurun_out/p11_b.txt
timeout 60 python tools/dbg_case.py 3,40,12,9 0 2>&1 | head -30 > gpurun_out/p11_c.txt
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_wd.so timeout 60 python tools/dbg_case.py 8,64,10,30 0 2>&1 | grep -v "^  " | head -60 > gpurun_out/p11_d.txt
