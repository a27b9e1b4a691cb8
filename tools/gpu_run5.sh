#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_probe5.py 2>&1 | grep -v -i warn | tee gpurun_out/r5_probe.log
echo "== engine tests"
timeout 900 python -m pytest tests/test_gpu_conv_engine.py -q -m gpu --maxfail=30 2>&1 | tail -30 | tee gpurun_out/r5_tests.log
echo "== all other gpu tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu --maxfail=30 2>&1 | tail -15 | tee gpurun_out/r5_tests2.log
echo "== breakdown"
for args in "--batch 512" "--batch 1024"; do
  timeout 300 python tools/step_breakdown.py $args 2>&1 | grep -v -i Warn | tail -16
done | tee gpurun_out/r5_breakdown.log
LPB_NO_IMPLICIT=1 timeout 300 python tools/step_breakdown.py --batch 512 2>&1 | grep -v -i Warn | tail -16 | tee -a gpurun_out/r5_breakdown.log
