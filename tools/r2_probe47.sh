#!/bin/bash
cd /root/repo
timeout 100 python -m pytest tests/test_reference_golden.py -m gpu -q 2>&1 | grep -v "Warning\|warnings" | cut -c1-300 | tail -60 > gpurun_out/p47.txt 2>&1
