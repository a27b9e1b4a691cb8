#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_graph_diag.py 2>&1 | grep -v "Warning\|run_backward" > gpurun_out/r2_22_diag.log
cat gpurun_out/r2_22_diag.log | cut -c1-300
