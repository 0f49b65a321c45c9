#!/bin/bash
# row helpers on wide rows: parity on the watchdog build first (a protocol bug traps instead of hanging), then on the regular
# build, then device time with and without them
cd /root/repo
D=tf_seq2seq_losses_b200
{
echo "== watchdog build"
CTCB200_LIB=$D/libctc_b200_wd.so timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "row_helpers" 2>&1 | tail -5
echo "== regular build"
timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "row_helpers or worker_configurations" 2>&1 | tail -5
for v in classic simplified; do
CTCB200_TVL=1600,5000,400 timeout 100 python tools/bsweep.py $v 256 2>&1 | grep "B="
CTCB200_TVL=1600,5000,400 CTCB200_PLAN=0,0,0,0,8 timeout 100 python tools/bsweep.py $v 256 2>&1 | grep "B="
CTCB200_TVL=1600,5000,400 CTCB200_PLAN=2,2,0,4,0 timeout 100 python tools/bsweep.py $v 256 2>&1 | grep "B="
done
} > gpurun_out/p43.txt 2>&1
