#!/usr/bin/env bash
# Round 2, run 18: complex64 instantiation of the samples-minor kernel (chain_wide.cuh<cplxf>).
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -x -q -k "complex64 or samples_minor" 2>&1 | tail -8 | tee $O/r2ac_tests.log
( timeout 300 python profiles/cfg5_probe.py cfg5 512 c64
  timeout 300 python profiles/cfg5_probe.py big4 1 c64 ) 2>&1 | tee $O/r2ac_c64_wide.txt
