#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/dbg_case.py 3,6,5,3 0 > gpurun_out/p9_dbg_a.txt 2>&1
timeout 60 python tools/dbg_case.py 3,6,8,3 0 > gpurun_out/p9_dbg_b.txt 2>&1
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_nowd.so timeout 60 python tools/dbg_case.py 3,6,5,3 0 > gpurun_out/p9_dbg_c.txt 2>&1
timeout 60 python tools/dbg_case.py 3,6,5,3 1 > gpurun_out/p9_dbg_d.txt 2>&1
