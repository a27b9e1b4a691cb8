#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r35_bench.log | cut -c1-200
