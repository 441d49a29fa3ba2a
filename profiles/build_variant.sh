#!/usr/bin/env bash
# Builds a variant of libqocvl.so whose chain_duo.cu is compiled with extra flags (kernel experiments):
#   profiles/build_variant.sh u2 -DDUO_UNROLL=2   ->  optimal-control-rydberg_b200/build/variants/libqocvl_u2.so
# (select it with QOCVL_LIB=...; the other objects come from the last build.sh run)
set -euo pipefail
PKG="$(cd "$(dirname "${BASH_SOURCE[0]}")/../optimal-control-rydberg_b200" && pwd)"
name="$1"; shift
mkdir -p "$PKG/build/variants"
# UNITS="chain_wide chain_wide_b13 chain_wide_b13f" profiles/build_variant.sh c4 -DWIDE_CHUNK=4   (other units than chain_duo)
UNITS="${UNITS:-chain_duo}"
objs=(); skip=" "
for u in $UNITS; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c "$PKG/csrc/$u.cu" -o "$PKG/build/variants/${u}_$name.o" &
  objs+=("$PKG/build/variants/${u}_$name.o"); skip="$skip$u.o "
done
wait
for o in "$PKG"/build/*.o; do case "$skip" in *" $(basename "$o") "*) ;; *) objs+=("$o");; esac; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$PKG/build/variants/libqocvl_$name.so" "${objs[@]}" -lcudart -ldl
echo "built $PKG/build/variants/libqocvl_$name.so"
