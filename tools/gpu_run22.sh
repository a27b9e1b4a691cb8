#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv_engine.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r22_tests.log
timeout 300 python tools/gpu_probe_pair.py 2>&1 | grep -v -i Warn | tee gpurun_out/r22_probe.log
timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -20 | tee gpurun_out/r22_breakdown_auto.log
LPB_GEMM_MODE=2 timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -20 | tee gpurun_out/r22_breakdown_persistent.log
LPB_GEMM_MODE=0 timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -20 | tee gpurun_out/r22_breakdown_single.log
