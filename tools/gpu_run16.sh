#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_probe16.py 2>&1 | grep -v -i warn | tee gpurun_out/r16_probe.log
echo "== full gpu tests"
timeout 1800 python -m pytest tests -q -m gpu --maxfail=30 2>&1 | tail -8 | tee gpurun_out/r16_tests.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke | tee gpurun_out/r16_smoke.log
echo "== bench default"
timeout 900 python bench.py --predictive 2>&1 | tail -1 | tee gpurun_out/r16_bench.log
