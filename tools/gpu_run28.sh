#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r28_tests.log
timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -22 | tee gpurun_out/r28_breakdown.log
LPB_NO_STRIDED=1 timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | head -3 | tee gpurun_out/r28_breakdown_nostrided.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r28_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
