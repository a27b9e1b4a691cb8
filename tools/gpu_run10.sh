#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r10_launches.csv \
  python tools/step_breakdown.py --batch 2048 > gpurun_out/r10_ncu.log 2>&1
echo "== ncu full: MN-major SYRK"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_tc_kernel<3, true>|gemm_nt_tc_kernelILi3ELb1" -s 117 -c 8 -o gpurun_out/r10_syrk -f \
  python tools/step_breakdown.py --batch 2048 > gpurun_out/r10_ncu_syrk.log 2>&1
echo "== ncu full: implicit conv"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_nhwc_tc_kernel -s 78 -c 6 -o gpurun_out/r10_conv -f \
  python tools/step_breakdown.py --batch 2048 > gpurun_out/r10_ncu_conv.log 2>&1
ls -la gpurun_out | grep r10; du -sh gpurun_out
