#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1300 -c 640 --csv --log-file gpurun_out/r2_11_launches.csv python tools/step_breakdown.py --batch 4096 --precision auto > gpurun_out/r2_11_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_nhwc_tc_persistent -s 60 -c 2 -o gpurun_out/r2_11_conv python tools/step_breakdown.py --batch 4096 --precision auto > gpurun_out/r2_11_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent_kernel -s 100 -c 3 -o gpurun_out/r2_11_syrk python tools/step_breakdown.py --batch 4096 --precision auto > gpurun_out/r2_11_ncu3.log 2>&1
ls -la gpurun_out | grep r2_11
