#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tests/diagnostics/gpu_smoke_prederr.py > gpurun_out/r2_3_smokeerr.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_3_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_3_bench.log 2>&1
timeout 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -k "frontend or quadform or conv_kron or eigh or jacobi" 2>&1 | tail -40 > gpurun_out/r2_3_tests.log
timeout 200 python tools/gpu_eigh_timing.py > gpurun_out/r2_3_eigh.log 2>&1
tail -3 gpurun_out/r2_3_tests.log
