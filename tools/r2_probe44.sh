#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -q -k "row_helpers and (3000 or 5000)" 2>&1 | grep -v "^  warnings\|Warning" | cut -c1-400 > gpurun_out/p44.txt 2>&1
