#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r32_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke | tee gpurun_out/r32_smoke.log
timeout 1200 python bench.py --predictive 2>&1 | tail -1 | tee gpurun_out/r32_bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r32_bench_ref.log | cut -c1-200
