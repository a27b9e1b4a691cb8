#!/usr/bin/env bash
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --model wrn28_10 --structure diag_ef --batch 256 --steps 10 --warmup 3 > gpurun_out/r2_17_wrn_n8.log 2>&1
tail -c 400 gpurun_out/r2_17_wrn_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --model vit_b16 --batch 32 --steps 5 --warmup 3 --no-predictive --no-cpu-baseline > gpurun_out/r2_17_vit_n8.log 2>&1
tail -c 400 gpurun_out/r2_17_vit_n8.log
