#!/usr/bin/env bash
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "reduced_system_kernels" > $O/r2n_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2n_quick.log
tail -12 $O/r2n_quick.log
for S in 512 1024 1536 2048; do
  timeout 300 python profiles/ensemble_probe.py $S "" "strong_chunks=4" "strong_chunks=8" "strong_chunks=16"
done 2>&1 | tee $O/r2n_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2n_launches_512.csv python profiles/ensemble_probe.py 512 "" > /dev/null 2>&1
grep -E "k_duo" $O/r2n_launches_512.csv | tail -8 | awk -F'","' '{print $5, $NF}'
