#!/usr/bin/env bash
# Round 2, run 21: cached evaluation graphs (api.cu::run_evaluation).
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r2ag_tests.log
timeout 300 python profiles/single_sample_probe.py 2>&1 | grep -v Warn | tee $O/r2ag_single.txt
timeout 300 python profiles/optimise_loop_timing.py 2>&1 | grep -v Warn | tail -5 | tee -a $O/r2ag_single.txt
