#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python bench.py --model wrn28_10 --structure diag_ef --batch 256 --steps 5 --warmup 3 > gpurun_out/r2_16_wrn_n1.log 2>&1
tail -c 1500 gpurun_out/r2_16_wrn_n1.log
timeout 900 python bench.py --model vit_b16 --batch 32 --steps 5 --warmup 3 --no-predictive --no-cpu-baseline > gpurun_out/r2_16_vit_n1.log 2>&1
tail -c 1500 gpurun_out/r2_16_vit_n1.log
