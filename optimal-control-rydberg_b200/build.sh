#!/usr/bin/env bash
# Builds libqocvl.so (the C-ABI library, include/qocvl.h) for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/libqocvl.so"
SRCS=("$HERE"/csrc/plan.cu "$HERE"/csrc/pulse_map.cu "$HERE"/csrc/chain_small.cu "$HERE"/csrc/chain_aug.cu "$HERE"/csrc/chain_big.cu "$HERE"/csrc/chain_setup.cu "$HERE"/csrc/dense.cu "$HERE"/csrc/api.cu)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared ${QOCVL_PTXAS_V:+-Xptxas -v} \
  -o "$OUT" "${SRCS[@]}" -lcudart
echo "built $OUT"
