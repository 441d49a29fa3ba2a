#!/usr/bin/env bash
# Round 2, GPU call 13: lean reverse sweep (no Xbar on entries only a dead pulse channel would consume), 3 CTAs / SM forward
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_reference_golden.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -m gpu > $O/r2r_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2r_quick.log
tail -5 $O/r2r_quick.log
for S in 4096 512; do
  timeout 300 python profiles/ensemble_probe.py $S "" "duo_nolean=1"
done 2>&1 | tee $O/r2r_probe.txt
for S in 4096 512; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2r_launches_$S.csv python profiles/ensemble_probe.py $S "" > /dev/null 2>&1
grep -E "k_duo" $O/r2r_launches_$S.csv | tail -8 | awk -F'","' -v s=$S '{print s, $5, $NF}'
done
