#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== targeted tests"
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv_engine.py -q -m gpu --maxfail=30 2>&1 | tail -8 | tee gpurun_out/r13_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config1 or zoo" 2>&1 | grep -E "passed|failed|Error|error" | tee -a gpurun_out/r13_tests.log
echo "== breakdown"
timeout 300 python tools/step_breakdown.py --batch 2048 2>&1 | grep -v -i Warn | tail -18 | tee gpurun_out/r13_breakdown.log
echo "== bench"
for b in 2048 4096 512; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $b 2>&1 | tail -1 | tee gpurun_out/r13_bench_$b.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print($b, d['value'], d['ms_per_step'], d['e2e']['value'])"
done
echo "== vit small"
timeout 600 python tools/step_breakdown.py --batch 64 --model vit_b16 2>&1 | grep -v -i Warn | tail -16 | tee gpurun_out/r13_vit.log
