#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -q -m gpu --maxfail=30 2>&1 | tail -12 | tee gpurun_out/r9_tests.log
echo "== bench default"
timeout 900 python bench.py --predictive 2>&1 | tail -1 | tee gpurun_out/r9_bench.log
timeout 900 python bench.py --batch 512 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r9_bench_b512.log
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r9_bench_ref.log
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r9_launches.csv \
  python tools/step_breakdown.py --batch 2048 > gpurun_out/r9_ncu.log 2>&1
echo "== ncu full SYRK (MN-major) + conv"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_tc_kernel|conv_nhwc_tc_kernel" -s 200 -c 40 -o gpurun_out/r9_tc -f \
  python tools/step_breakdown.py --batch 2048 > gpurun_out/r9_ncu_full.log 2>&1
ls -la gpurun_out | grep r9
