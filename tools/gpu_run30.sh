#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_engine.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r30_tests.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "SYRK_ROTATE=$1 CONV_ROTATE=$2" | tee -a gpurun_out/r30_ab.log
  LPB_SYRK_ROTATE=$1 LPB_CONV_ROTATE=$2 timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -E "ms/step device|conv_nhwc|syrk_conv|conv_bwd_strided" | tee -a gpurun_out/r30_ab.log
done
