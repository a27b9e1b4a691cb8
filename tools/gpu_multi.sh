#!/usr/bin/env bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi_bench_n$N.log 2>&1
tail -c 600 gpurun_out/multi_bench_n$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py > gpurun_out/multi_dist_n$N.log 2>&1
tail -5 gpurun_out/multi_dist_n$N.log
