#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
  tools/dist_check.py 2>&1 | grep dist_check | tee gpurun_out/r33_dist_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r33_bench_n2.log | cut -c1-300
