#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Warning\|^  " | tail -12 > gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_ref.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench.log 2>&1
tail -3 gpurun_out/r2_final_tests.log; tail -2 gpurun_out/r2_final_smoke.log; tail -c 300 gpurun_out/r2_final_ref.log
