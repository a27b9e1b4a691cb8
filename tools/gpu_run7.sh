#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 1500 python -m pytest tests -q -m gpu --maxfail=30 2>&1 | tail -25 | tee gpurun_out/r7_tests.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r7_smoke.log
echo "== breakdown"
for args in "--batch 512" "--batch 2048"; do
  timeout 300 python tools/step_breakdown.py $args 2>&1 | grep -v -i Warn | tail -16
done | tee gpurun_out/r7_breakdown.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r7_bench.log
