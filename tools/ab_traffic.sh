#!/bin/bash
# Developer tool: DRAM traffic and duration of the fused kernel for each library variant given (suffixes of libctc_b200*.so)
for v in "$@"; do
  [ "$v" = "default" ] && v=""
  echo "== variant '$v'"
  CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200$v.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:kf_fused -s 3 -c 1 python tools/ab.py simple 2>&1 | grep -E "dram__|gpu__time|lts__"
done
