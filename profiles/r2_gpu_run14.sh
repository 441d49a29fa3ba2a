#!/usr/bin/env bash
# Round 2, GPU call 16: next slice prefetched (forward: second register set; reverse: into the dead entry registers before the Xbar staging)
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_reference_golden.py -x -q -m gpu -k "reduced_system_kernels or cfg3" > $O/r2x_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2x_quick.log
tail -3 $O/r2x_quick.log
timeout 300 python profiles/ensemble_probe.py 4096 "" 2>&1 | tee $O/r2x_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2x_launches.csv python profiles/ensemble_probe.py 4096 "" > /dev/null 2>&1
grep -E "k_duo" $O/r2x_launches.csv | tail -5 | awk -F'","' '{print $5, $NF}'
