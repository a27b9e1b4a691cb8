#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_engine.py tests/test_gpu_parity.py -q -m gpu --maxfail=30 2>&1 | tail -6 | tee gpurun_out/r17_tests.log
timeout 300 python tools/step_breakdown.py --batch 2048 2>&1 | grep -v -i Warn | tail -18 | tee gpurun_out/r17_breakdown.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --predictive 2>&1 | tail -1 | tee gpurun_out/r17_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('glm_predictive_ll_full_samples_per_sec'))"
timeout 900 python tools/gpu_probe17.py 2>&1 | grep -v -i warn | tee gpurun_out/r17_wrn.log
