#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== ncu launches (implicit engine)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r6_launches.csv \
  python tools/step_breakdown.py --batch 512 > gpurun_out/r6_ncu.log 2>&1
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 --predictive 2>&1 | tail -1 | tee gpurun_out/r6_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --batch 2048 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r6_bench_b2048.log
