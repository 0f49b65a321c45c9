#!/bin/bash
# K4 with the compacted work list (dynamic (token, direction) items) + which round-2 commit made the wide-row shape slower
cd /root/repo
D=tf_seq2seq_losses_b200
{
timeout 400 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "hessian or hvp or readme or second_derivative" 2>&1 | tail -5
python tools/bench_configs.py hessian
CTCB200_LIB=$D/libctc_b200_noscat.so python tools/bench_configs.py hessian
L="$D/libctc_b200.so $D/libctc_b200_r1.so $D/libctc_b200_b1312c37.so $D/libctc_b200_b8a95621.so $D/libctc_b200_bbe747c8.so $D/libctc_b200_be99eca0.so"
python tools/ab_lib.py classic 256,1600,5000,400 $L
python tools/ab_lib.py simplified 256,1600,5000,400 $L
python tools/ab_lib.py simplified 256,1000,1024,200 $L
} > gpurun_out/p42.txt 2>&1
