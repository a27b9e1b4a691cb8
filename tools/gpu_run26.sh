#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r26_tests.log
timeout 900 python tools/gpu_probe17.py 2>&1 | grep -v -i warn | tee gpurun_out/r26_wrn.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r26_launches.csv \
  python tools/step_breakdown.py --batch 4096 > gpurun_out/r26_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_nhwc_tc_persistent_kernel" -s 30 -c 4 -o gpurun_out/r26_conv -f \
  python tools/step_breakdown.py --batch 4096 > gpurun_out/r26_ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_persistent_kernel" -s 60 -c 8 -o gpurun_out/r26_gemm -f \
  python tools/step_breakdown.py --batch 4096 > gpurun_out/r26_ncu_gemm.log 2>&1
ls -la gpurun_out | grep r26
timeout 1200 python bench.py --predictive 2>&1 | tail -1 | tee gpurun_out/r26_bench.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r26_bench_ref.log | cut -c1-400
