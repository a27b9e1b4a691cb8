#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cuda_graph" 2>&1 | grep -v "Warning" | tail -60 > gpurun_out/r2_21_tests.log
tail -3 gpurun_out/r2_21_tests.log
