#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_decompose_profile.py > gpurun_out/r2_8_decomp.log 2>&1
cat gpurun_out/r2_8_decomp.log | grep -v Warn | tail -20
