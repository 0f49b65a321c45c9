#!/bin/bash
mkdir -p gpurun_out
for plan in "nohalf W4" "split W8 nohalf" "half W4" "split W8" "W2 R4" "W1 R1" "default"; do
  echo "===== plan $plan (watchdog lib)" >> gpurun_out/p18_quick.txt
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_wdog.so timeout 100 python tools/quickcheck.py "$plan" >> gpurun_out/p18_quick.txt 2>&1
done
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_wdog.so timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 --timeout=300 > gpurun_out/p18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p18_pytest.log
timeout 200 python tools/bsweep.py simplified 256,128,64,32 > gpurun_out/p18_bsweep_simple.txt 2>&1
timeout 200 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/p18_bsweep_classic.txt 2>&1
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_norec.so timeout 200 python tools/bsweep.py simplified 256,32 > gpurun_out/p18_bsweep_simple_norec.txt 2>&1
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_norec.so timeout 200 python tools/bsweep.py classic 256,32 > gpurun_out/p18_bsweep_classic_norec.txt 2>&1
