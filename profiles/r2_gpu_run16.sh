#!/usr/bin/env bash
# Round 2, GPU call 18: three basis columns per item in the chunk-propagator sweep
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_reference_golden.py -x -q -m gpu -k "reduced_system_kernels or cfg3 or other_small or complex64" > $O/r2aa_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2aa_quick.log
tail -5 $O/r2aa_quick.log
for S in 512 1024 2048; do timeout 300 python profiles/ensemble_probe.py $S "" "duo_cols=0"; done 2>&1 | tee $O/r2aa_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2aa_launches_512.csv python profiles/ensemble_probe.py 512 "" > /dev/null 2>&1
grep -E "k_duo" $O/r2aa_launches_512.csv | tail -8 | awk -F'","' '{print $5, $NF}'
cd profiles && python single_sample_probe.py "" | tail -1
