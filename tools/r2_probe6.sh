#!/bin/bash
set -x
mkdir -p gpurun_out
# the test that hung in probe 5, repeated, with the verbose watchdog build
for i in 1 2 3 4 5 6; do
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_wd.so timeout 120 python -m pytest tests/test_cuda_parity.py -m gpu -q -x --timeout=100 -k "test_loss_and_gradient_match_oracle and fused" >> gpurun_out/p6_repeat.log 2>&1
  echo "iteration $i rc=$?" >> gpurun_out/p6_repeat.log
done
timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 --timeout=300 > gpurun_out/p6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p6_pytest.log
timeout 200 python tools/bsweep.py simplified 256,128,64,32 > gpurun_out/p6_bsweep_simple.txt 2>&1
timeout 200 python tools/bsweep.py classic 256,128,64,32 > gpurun_out/p6_bsweep_classic.txt 2>&1
