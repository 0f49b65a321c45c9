#!/bin/bash
# K4 compiled for four resident CTAs per SM (64 registers) against the default; ncu --set full of the wide-row fused launch
cd /root/repo
D=tf_seq2seq_losses_b200
{
python tools/bench_configs.py hessian
CTCB200_LIB=$D/libctc_b200_k4b4.so python tools/bench_configs.py hessian
python tools/bench_configs.py hessian
CTCB200_LIB=$D/libctc_b200_k4b4.so python tools/bench_configs.py hessian
} > gpurun_out/p45.txt 2>&1
CTCB200_TVL=1600,5000,400 timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:kf_fused -s 3 -c 1 -o gpurun_out/r2b_fused_classic_cfg4 python tools/bsweep.py classic 256 >> gpurun_out/p45.txt 2>&1
