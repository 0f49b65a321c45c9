#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/p29_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p29_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/p29_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/p29_bench.json 2> gpurun_out/p29_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/p29_bench_ref.json 2>> gpurun_out/p29_bench.err
