#!/usr/bin/env bash
# Builds the two sm_100a microbenchmarks next to their sources (run them on a B200: ./dfma_operands ; ./latency).
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")"
for n in dfma_operands latency half_warp; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o $n $n.cu
done
