#!/usr/bin/env bash
# Builds libqocvl.so (the C-ABI library, include/qocvl.h) for sm_100a, in-tree.
# Every translation unit is compiled on its own (in parallel), then linked; no relocatable device code needed.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/libqocvl.so"
OBJ="$HERE/build"
mkdir -p "$OBJ"
UNITS=(plan pulse_map chain_small chain_aug chain_duo chain_big chain_wide chain_wide_b13 chain_wide_b24 chain_wide_b13f chain_wide_b24f chain_setup dense dense_blocks api lindblad)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${QOCVL_PTXAS_V:+-Xptxas -v})
pids=()
for u in "${UNITS[@]}"; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/csrc/$u.cu" -o "$OBJ/$u.o" > "$OBJ/$u.log" 2>&1 &
  pids+=($!)
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then fail=1; echo "nvcc failed on ${UNITS[$i]}.cu:" >&2; fi
  cat "$OBJ/${UNITS[$i]}.log" >&2
done
[ "$fail" -eq 0 ] || exit 1
objs=()
for u in "${UNITS[@]}"; do objs+=("$OBJ/$u.o"); done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$OUT" "${objs[@]}" -lcudart -ldl
echo "built $OUT"
