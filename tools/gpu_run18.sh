#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=30 2>&1 | tail -6 | tee gpurun_out/r18_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke | tee gpurun_out/r18_smoke.log
timeout 900 python bench.py --predictive 2>&1 | tail -1 | tee gpurun_out/r18_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v for k,v in d['config'].items() if 'per_sec' in k or 'ms' in k}, d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r18_bench_ref.log | cut -c1-200
