#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r11_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | grep -v -i warn | tail -5 | tee gpurun_out/r11_dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r11_bench_n2.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r11_bench_n1.log
