#!/usr/bin/env bash
# 8-GPU check: sharded fit == single-process fit (dist_check) and the scaling bench line at N = 8 / 4 / 2.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r25_gpus.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  tools/dist_check.py 2>&1 | grep -v -i warn | tail -6 | tee gpurun_out/r25_dist_check.log
for n in 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
  bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r25_bench_n$n.log | cut -c1-300
done
