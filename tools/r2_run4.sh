#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_parity.py -m gpu -q -k "resnet_through or gridsearch or conv_kron_predictive_without" 2>&1 | grep -v "^  warnings\|Warning" | tail -150 > gpurun_out/r2_4_tests.log
