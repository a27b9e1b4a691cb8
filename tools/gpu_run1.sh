#!/usr/bin/env bash
# First GPU bring-up: SIMT kernels, then tcgen05 kernel under a timeout, then parity, smoke, bench.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0))" > gpurun_out/env.txt 2>&1
echo "== kernels (no tensor core)" 
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not tensor_core" --maxfail=50 2>&1 | tail -40 | tee gpurun_out/t1_kernels.log
echo "== tensor core gemm"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tensor_core" --maxfail=50 2>&1 | tail -60 | tee gpurun_out/t2_tc.log
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --maxfail=60 2>&1 | tail -80 | tee gpurun_out/t3_parity.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/t4_smoke.log
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 --predictive 2>&1 | tail -5 | tee gpurun_out/t5_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision bf16 --no-cpu-baseline 2>&1 | tail -5 | tee gpurun_out/t6_bench_bf16.log
