#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 1500 python -m pytest tests -q -m gpu --maxfail=30 2>&1 | tail -25 | tee gpurun_out/r8_tests.log
echo "== breakdown"
for args in "--batch 512" "--batch 2048"; do
  timeout 300 python tools/step_breakdown.py $args 2>&1 | grep -v -i Warn | tail -17
done | tee gpurun_out/r8_breakdown.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 --predictive 2>&1 | tail -1 | tee gpurun_out/r8_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --batch 2048 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r8_bench_b2048.log
