#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "decompose_routes or eigh" 2>&1 | grep -v "Warning\|^  " | tail -30 > gpurun_out/r2_10_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_10_smoke.log 2>&1
timeout 300 python tools/gpu_decompose_profile.py > gpurun_out/r2_10_decomp.log 2>&1
timeout 300 python tests/diagnostics/gpu_smoke_prederr.py > gpurun_out/r2_10_smokeerr.log 2>&1
tail -3 gpurun_out/r2_10_tests.log; tail -3 gpurun_out/r2_10_smoke.log; grep "whole\|library" gpurun_out/r2_10_decomp.log
