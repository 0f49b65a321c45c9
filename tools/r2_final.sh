#!/bin/bash
# Round-2 final validation on one B200: GPU test suite, smoke, both bench arms, batch sweep of both variants.
cd /root/repo
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference 2>&1 | tail -1
python bench.py 2>&1 | tail -1
python bench.py --variant classic --no-cpu-baseline 2>&1 | tail -1
python tools/bsweep.py classic 256 128 64 32 2>&1 | grep kf_
python tools/bsweep.py simplified 256 128 64 32 2>&1 | grep kf_
} > gpurun_out/final.txt 2>&1
