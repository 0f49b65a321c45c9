#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/p30_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p30_pytest.log
