#!/usr/bin/env bash
# Builds gp_grief_b200/_lib/libgrief_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../_lib"
mkdir -p "${OUT}" "${HERE}/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
       --expt-relaxed-constexpr -I"${HERE}/../../include")
SRCS=(plan rows phi_stage solve dense ozaki topk grad krmatvec comm abi)
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
  "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${s}.cu" -o "${HERE}/build/${s}.o" > "${HERE}/build/${s}.log" 2>&1 &
  pids+=($!)
  OBJS+=("${HERE}/build/${s}.o")
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then echo "compile failed: ${SRCS[$i]}"; cat "${HERE}/build/${SRCS[$i]}.log"; fail=1; fi
done
[ "${fail}" = 0 ] || exit 1
"${NVCC}" -shared -o "${OUT}/libgrief_b200.so" "${OBJS[@]}" -gencode arch=compute_100a,code=sm_100a \
  -L/usr/local/cuda/lib64 -lcudart -lcuda -ldl -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo "built ${OUT}/libgrief_b200.so"
