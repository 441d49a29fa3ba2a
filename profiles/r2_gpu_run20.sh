#!/usr/bin/env bash
# Round 2, run 20: k_wide_degrees (phase shift + 2-norm degrees on the samples-minor kernel).
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -12 | tee $O/r2ae_tests.log
( timeout 300 python profiles/cfg5_probe.py cfg5 512 c64
  timeout 300 python profiles/cfg5_probe.py big4 1 c64 ) 2>&1 | tee $O/r2ae_wide.txt
