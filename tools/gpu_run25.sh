#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_probe_scale.py 2>&1 | grep -v -i Warn | tee gpurun_out/r25_probe_scale.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r25_tests.log
timeout 300 python tools/step_breakdown.py --batch 512 2>&1 | grep -v -i Warn | head -3 | tee gpurun_out/r25_breakdown_512.log
timeout 300 python tools/step_breakdown.py --batch 2048 2>&1 | grep -v -i Warn | head -3 | tee gpurun_out/r25_breakdown_2048.log
