#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/p21_quick.txt
for plan in "nohalf W4" "split W8 nohalf" "W2 R4" "W1 R1" "default"; do
  echo "===== plan $plan (watchdog lib)" >> gpurun_out/p21_quick.txt
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_wdog.so timeout 100 python tools/quickcheck.py "$plan" >> gpurun_out/p21_quick.txt 2>&1
done
CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_wdog.so timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 --timeout=300 > gpurun_out/p21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p21_pytest.log
timeout 200 python tools/bsweep.py simplified 256,128,64,32 > gpurun_out/p21_bsweep_simple.txt 2>&1
timeout 200 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/p21_bsweep_classic.txt 2>&1
timeout 600 python bench.py > gpurun_out/p21_bench.json 2> gpurun_out/p21_bench.err
timeout 300 python bench.py --variant classic --no-e2e --no-cpu-baseline > gpurun_out/p21_bench_classic.json 2>> gpurun_out/p21_bench.err
