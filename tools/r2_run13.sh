#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv_engine.py -m gpu -q -x 2>&1 | grep -v "Warning\|^  " | tail -15 > gpurun_out/r2_13_tests.log
timeout 300 python tools/gpu_eigh_timing.py 2>&1 | grep -i "jacobi" > gpurun_out/r2_13_jacobi.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-predictive > gpurun_out/r2_13_bench.log 2>&1
tail -3 gpurun_out/r2_13_tests.log; cat gpurun_out/r2_13_jacobi.log
