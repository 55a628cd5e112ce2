#!/usr/bin/env bash
# Builds a variant library for kernel experiments: build_variant.sh <name> [extra nvcc flags...]
# -> vittf_b200/libvittf_b200_<name>.so (select with VITTF_LIB=<path>); only attention.cu / gemm.cu / similarity.cu
# are recompiled with the extra flags, the other objects are reused from build/.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NAME="$1"; shift
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr)
mkdir -p "${HERE}/build/${NAME}"
for f in attention gemm similarity; do
  "${NVCC}" "${FLAGS[@]}" "$@" -c "${HERE}/${f}.cu" -o "${HERE}/build/${NAME}/${f}.o" &
done
wait
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${HERE}/../libvittf_b200_${NAME}.so" \
  "${HERE}"/build/${NAME}/{attention,gemm,similarity}.o "${HERE}"/build/{api,vit_ops,vit_engine,sim_up_tc,bls,sampling}.o -lcudart
echo "built libvittf_b200_${NAME}.so"
