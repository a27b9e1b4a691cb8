#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_probe15.py 2>&1 | grep -v -i warn | tee gpurun_out/r15_probe.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "maxpool or elementwise" 2>&1 | tail -3 | tee gpurun_out/r15_tests.log
timeout 300 python tools/step_breakdown.py --batch 2048 2>&1 | grep -v -i Warn | tail -18 | tee gpurun_out/r15_breakdown.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_nt_tc_kernel<\(int\)3, \(bool\)1>" -s 117 -c 10 -o gpurun_out/r15_syrk -f \
  python tools/step_breakdown.py --batch 2048 > gpurun_out/r15_ncu_syrk.log 2>&1
ls -la gpurun_out | grep r15
