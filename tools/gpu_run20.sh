#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "cta_pair" 2>&1 | tail -15 | tee gpurun_out/r20_pair_tests.log
timeout 300 python tools/gpu_probe_pair.py 2>&1 | grep -v -i Warn | tee gpurun_out/r20_pair_probe.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r20_kernel_tests.log
LPB_GEMM_PAIR=0 timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -18 | tee gpurun_out/r20_breakdown_single.log
timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -18 | tee gpurun_out/r20_breakdown_pair.log
