#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_syev_batched_probe.py > gpurun_out/r2_9_syevb.log 2>&1
tail -12 gpurun_out/r2_9_syevb.log
