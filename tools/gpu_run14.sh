#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== targeted tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "maxpool or elementwise" 2>&1 | tail -3 | tee gpurun_out/r14_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config1" 2>&1 | grep -E "passed|failed|Error|error" | tee -a gpurun_out/r14_tests.log
echo "== breakdown"
timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -18 | tee gpurun_out/r14_breakdown.log
echo "== ncu full: MN-major SYRK"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_nt_tc_kernel<3, 1>" -s 117 -c 10 -o gpurun_out/r14_syrk -f \
  python tools/step_breakdown.py --batch 2048 > gpurun_out/r14_ncu_syrk.log 2>&1
ls -la gpurun_out | grep r14; du -sh gpurun_out
