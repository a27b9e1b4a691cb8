#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r27_tests.log
timeout 900 python tools/gpu_probe17.py 2>&1 | grep -v -i warn | tee gpurun_out/r27_wrn.log
timeout 300 python tools/step_breakdown.py --model vit_b16 --batch 64 2>&1 | grep -v -i Warn | tail -16 | tee gpurun_out/r27_vit_breakdown.log
