#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_probe4.py 2>&1 | grep -v -i warn | tee gpurun_out/r4_probe.log
echo "== ncu launches (engine)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r4_launches.csv \
  python tools/step_breakdown.py --batch 512 > gpurun_out/r4_ncu.log 2>&1
echo "== ncu full tc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc_kernel -s 250 -c 12 -o gpurun_out/r4_tc -f \
  python tools/step_breakdown.py --batch 512 > gpurun_out/r4_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
