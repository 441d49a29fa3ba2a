#!/usr/bin/env bash
# Round 2, run 17: per-role Taylor degrees on the two-lane kernels (roles = independent diagonal blocks).
O=gpurun_out
mkdir -p $O
( timeout 300 python profiles/ensemble_probe.py 4096 "" "duo_roleq=1" "duo_roleq=0"
  timeout 300 python profiles/ensemble_probe.py 512 "" "duo_roleq=0"
  timeout 300 python profiles/single_sample_probe.py ) > $O/r2ab_roleq.txt 2>&1
cat $O/r2ab_roleq.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r2ab_tests.log
