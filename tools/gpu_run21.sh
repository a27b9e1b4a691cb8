#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe_pair.py 2>&1 | grep -v -i Warn | tee gpurun_out/r21_pair_probe.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r21_tests.log
timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -20 | tee gpurun_out/r21_breakdown.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r21_launches.csv \
  python tools/step_breakdown.py --batch 4096 > gpurun_out/r21_ncu.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r21_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
