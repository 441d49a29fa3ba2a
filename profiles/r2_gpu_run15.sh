#!/usr/bin/env bash
# Round 2, GPU call 17: float2 instantiation of the two-lane kernels (complex64 mode)
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 900 > $O/r2y_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2y_tests.log
tail -12 $O/r2y_tests.log
timeout 300 python profiles/cfg3_probe.py c64 4096 2>&1 | tail -3 | tee $O/r2y_c64.txt
timeout 300 python profiles/ensemble_probe.py 4096 "" 2>&1 | tail -2
