#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== engine tests"
timeout 900 python -m pytest tests/test_gpu_conv_engine.py tests/test_gpu_kernels.py -q -m gpu --maxfail=30 2>&1 | tail -30 | tee gpurun_out/r3_tests.log
echo "== parity tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --maxfail=30 2>&1 | tail -30 | tee gpurun_out/r3_parity.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r3_smoke.log
echo "== breakdown"
for args in "--batch 512" "--batch 512 --no-engine" "--batch 1024" "--batch 2048"; do
  timeout 300 python tools/step_breakdown.py $args 2>&1 | grep -v -i Warn | tail -14
done | tee gpurun_out/r3_breakdown.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 --predictive 2>&1 | tail -2 | tee gpurun_out/r3_bench.log
