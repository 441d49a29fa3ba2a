#!/usr/bin/env bash
# Round 2, GPU call 6: operand-sharing order of the DFMAs; 4096 vs 8192 samples (two warps per scheduler in the forward kernel)
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "reduced_system_kernels" > $O/r2g_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2g_quick.log
for S in 4096 8192; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2g_launches_$S.csv python profiles/ensemble_probe.py $S "" > /dev/null 2>&1
  grep -E "k_duo_(forward|reverse)" $O/r2g_launches_$S.csv | tail -2 | awk -F'","' -v v=$S '{print v, $5, $NF}'
done
timeout 300 python profiles/ensemble_probe.py 4096 "" > $O/r2g_probe.txt 2>&1; cat $O/r2g_probe.txt
