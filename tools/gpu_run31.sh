#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r31_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke | tee gpurun_out/r31_smoke.log
LPB_NO_OVERLAP=1 timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | head -3 | tee gpurun_out/r31_breakdown_nooverlap.log
timeout 300 python tools/step_breakdown.py --batch 4096 2>&1 | grep -v -i Warn | tail -22 | tee gpurun_out/r31_breakdown.log
timeout 300 python tools/gpu_probe_scale.py 2>&1 | grep -v -i Warn | head -8 | tee gpurun_out/r31_probe_scale.log
timeout 1200 python bench.py 2>&1 | tail -1 | tee gpurun_out/r31_bench.log | cut -c1-300
