#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 1200 python -m pytest tests -q -m gpu --maxfail=30 2>&1 | tail -30 | tee gpurun_out/r2_tests.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2_smoke.log
echo "== breakdown"
for args in "--batch 512" "--batch 512 --tf32" "--batch 512 --loop-backward" "--batch 2048" "--batch 512 --precision bf16" "--batch 512 --precision fp32"; do
  timeout 300 python tools/step_breakdown.py $args 2>&1 | grep -v Warning | tail -9
done | tee gpurun_out/r2_breakdown.log
echo "== cpu thread sweep"
for t in 8 16 32 64; do
  timeout 300 python bench.py --impl reference --steps 1 --warmup 1 --cpu-threads $t 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($t, d['value'])"
done | tee gpurun_out/r2_cpu_threads.log
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv \
  python tools/step_breakdown.py --batch 512 > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
