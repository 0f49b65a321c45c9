#!/bin/bash
mkdir -p gpurun_out
for plan in "" 4,2,1,8,1 4,3,1,8,1 8,2,0,16,5 6,2,1,12,1 6,2,0,12,5; do
  echo "== plan '$plan'" >> gpurun_out/p26_b128.txt
  CTCB200_PLAN=$plan timeout 100 python tools/bsweep.py simplified 128,96 2>&1 | grep "B=" >> gpurun_out/p26_b128.txt
done
