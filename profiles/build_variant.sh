#!/usr/bin/env bash
# Builds a variant of libqocvl.so whose chain_duo.cu is compiled with extra flags (kernel experiments):
#   profiles/build_variant.sh u2 -DDUO_UNROLL=2   ->  optimal-control-rydberg_b200/build/variants/libqocvl_u2.so
# (select it with QOCVL_LIB=...; the other objects come from the last build.sh run)
set -euo pipefail
PKG="$(cd "$(dirname "${BASH_SOURCE[0]}")/../optimal-control-rydberg_b200" && pwd)"
name="$1"; shift
mkdir -p "$PKG/build/variants"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c "$PKG/csrc/chain_duo.cu" -o "$PKG/build/variants/chain_duo_$name.o"
objs=()
for o in "$PKG"/build/*.o; do [ "$(basename "$o")" = chain_duo.o ] || objs+=("$o"); done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$PKG/build/variants/libqocvl_$name.so" "${objs[@]}" "$PKG/build/variants/chain_duo_$name.o" -lcudart -ldl
echo "built $PKG/build/variants/libqocvl_$name.so"
