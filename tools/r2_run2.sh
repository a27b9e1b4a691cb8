#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_2_smoke.log 2>&1
timeout 400 python tools/gpu_forward_err.py > gpurun_out/r2_2_fwderr.log 2>&1
timeout 400 python tools/gpu_eigh_timing.py > gpurun_out/r2_2_eigh.log 2>&1
tail -5 gpurun_out/r2_2_tests.log
