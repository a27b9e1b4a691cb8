#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_1_smi.txt 2>&1
timeout 600 python tests/diagnostics/gpu_bisect_kfac.py > gpurun_out/r2_1_bisect.log 2>&1
timeout 400 python tools/gpu_eigh_timing.py > gpurun_out/r2_1_eigh.log 2>&1
timeout 400 python tests/diagnostics/gpu_predictive_err.py > gpurun_out/r2_1_prederr.log 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_1_tests.log
tail -5 gpurun_out/r2_1_tests.log
