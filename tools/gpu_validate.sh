#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Warning\|^  " | tail -40 > gpurun_out/val_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/val_bench.log 2>&1
tail -3 gpurun_out/val_tests.log
