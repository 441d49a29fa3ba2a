#!/usr/bin/env bash
# Round 2, GPU call 7: re / im lane split (4 lanes per sector, two warps per scheduler)
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_reference_golden.py -x -q -m gpu -k "reduced_system_kernels or cfg3_full_size_members" > $O/r2h_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2h_quick.log
tail -5 $O/r2h_quick.log
for S in 4096 512; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2h_launches_$S.csv python profiles/ensemble_probe.py $S "" > /dev/null 2>&1
  grep -E "k_duo_(forward|reverse)" $O/r2h_launches_$S.csv | tail -2 | awk -F'","' -v v=$S '{print v, $5, $NF}'
done
timeout 300 python profiles/ensemble_probe.py 4096 "" "duo_wpc=4" > $O/r2h_probe.txt 2>&1; cat $O/r2h_probe.txt
