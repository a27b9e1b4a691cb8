#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -q -m gpu --maxfail=30 2>&1 | tail -12 | tee gpurun_out/r12_tests.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke | tee gpurun_out/r12_smoke.log
echo "== breakdown"
timeout 300 python tools/step_breakdown.py --batch 2048 2>&1 | grep -v -i Warn | tail -18 | tee gpurun_out/r12_breakdown.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r12_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 512 2>&1 | tail -1 | tee gpurun_out/r12_bench_512.log
