#!/usr/bin/env bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none -k regex:"conv_nhwc_tc_persistent|gemm_tc_persistent|kron_conv_quadform" -s 520 -c 135 --csv --log-file gpurun_out/ncu_tensor_kernels.csv python tools/step_breakdown.py --batch 4096 --precision auto > gpurun_out/ncu_tensor_kernels.log 2>&1
wc -l gpurun_out/ncu_tensor_kernels.csv
