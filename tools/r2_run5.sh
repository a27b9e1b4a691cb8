#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tests/diagnostics/gpu_lean_diag.py > gpurun_out/r2_5_lean.log 2>&1
timeout 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_parity.py tests/test_gpu_conv_engine.py -m gpu -q -k "frontend or conv_kron or gp_kernels or mc_fisher or two_product" 2>&1 | grep -v "Warning\|^  " | tail -60 > gpurun_out/r2_5_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_5_bench.log 2>&1
tail -3 gpurun_out/r2_5_tests.log
