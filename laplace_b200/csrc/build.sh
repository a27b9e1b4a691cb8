#!/usr/bin/env bash
# Build liblaplace_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../lib"
mkdir -p "${OUT}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --use_fast_math=false)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC)
SRCS=(api.cu pack.cu gemm_simt.cu gemm_tc.cu gemm_tc2.cu gemm_tc3.cu conv_tc.cu misc.cu eigh.cu elementwise.cu kronquad.cu)
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
  o="${OUT}/${s%.cu}.o"
  OBJS+=("$o")
  "${NVCC}" "${FLAGS[@]}" ${LPB_PTXAS_V:+-Xptxas -v} -c "${HERE}/${s}" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"${NVCC}" -shared -o "${OUT}/liblaplace_b200.so" "${OBJS[@]}" -cudart static
echo "built ${OUT}/liblaplace_b200.so"
