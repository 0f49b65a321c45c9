#!/bin/bash
# Final validation after the helpers became a compile-time switch: GPU suite, smoke, headline + classic bench lines, A/B against
# the library of the commit before the helpers, the wide-row shape.
cd /root/repo
D=tf_seq2seq_losses_b200
{
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 60 python tools/ab_lib.py classic 256,1000,1024,200 $D/libctc_b200.so $D/libctc_b200_pre.so
timeout 60 python tools/ab_lib.py classic 32,1000,1024,200 $D/libctc_b200.so $D/libctc_b200_pre.so
CTCB200_TVL=1600,5000,400 timeout 60 python tools/bsweep.py classic 256 2>&1 | grep "B="
CTCB200_TVL=1600,5000,400 timeout 60 python tools/bsweep.py simplified 256 2>&1 | grep "B="
} > gpurun_out/final3.txt 2>&1
timeout 200 python bench.py 2>/dev/null | tail -1 > gpurun_out/r2_bench_default.json
timeout 100 python bench.py --variant classic --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r2_bench_classic.json
