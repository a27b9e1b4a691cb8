#!/usr/bin/env bash
mkdir -p gpurun_out
for pf in 0 1; do
echo "LPB_NO_PREFETCH=$pf" | tee -a gpurun_out/r34_n2.log
LPB_NO_PREFETCH=$pf timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$pf \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2))" | tee -a gpurun_out/r34_n2.log
done
