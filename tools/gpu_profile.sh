#!/usr/bin/env bash
# ncu evidence for profiles/: (1) launch list with device times of one bench command, (2) full capture of the
# tcgen05 SYRK kernel.  Never used for bench numbers.
mkdir -p gpurun_out
PREC="${1:-bf16x3}"
BATCH="${2:-512}"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${PREC}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision ${PREC} --batch ${BATCH} > gpurun_out/prof_bench_${PREC}.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc_kernel -s 60 -c 6 -o gpurun_out/prof_tc_${PREC} -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision ${PREC} --batch ${BATCH} > gpurun_out/prof_full_${PREC}.log 2>&1
ls -la gpurun_out
