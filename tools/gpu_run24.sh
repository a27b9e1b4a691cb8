#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r24_tests.log
LPB_CONV_MODE=0 timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -22 | tee gpurun_out/r24_breakdown_conv_single.log
timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -22 | tee gpurun_out/r24_breakdown.log
timeout 300 python tools/step_breakdown.py --batch 512 2>&1 | grep -v -i Warn | head -3 | tee gpurun_out/r24_breakdown_512.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r24_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
