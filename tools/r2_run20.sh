#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frontend.py tests/test_gpu_kernels.py -m gpu -q -k "cuda_graph or frontend or maxpool" 2>&1 | grep -v "Warning\|^  " | tail -40 > gpurun_out/r2_20_tests.log
timeout 600 python bench.py > gpurun_out/r2_20_bench.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-predictive > gpurun_out/r2_20_bench2.log 2>&1
tail -3 gpurun_out/r2_20_tests.log
