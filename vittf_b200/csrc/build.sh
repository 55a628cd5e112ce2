#!/usr/bin/env bash
# Builds libvittf_b200.so (sm_100a only) next to the Python package.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libvittf_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
       --expt-relaxed-constexpr -Xptxas -v)
mkdir -p "${HERE}/build"
pids=()
for f in api gemm attention vit_ops vit_engine similarity sim_up_tc bls sampling; do
  ( "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/build/${f}.o" > "${HERE}/build/${f}.log" 2>&1 \
      || { cat "${HERE}/build/${f}.log"; exit 1; } ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
[ $rc -eq 0 ] || { echo "build failed"; exit 1; }
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${OUT}" "${HERE}"/build/{api,gemm,attention,vit_ops,vit_engine,similarity,sim_up_tc,bls,sampling}.o -lcudart
echo "built ${OUT}"
