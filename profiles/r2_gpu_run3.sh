#!/usr/bin/env bash
# Round 2, GPU call 3: first run of the two-lane kernels (chain_duo.cu): agreement tests, timings, launch list.
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -k "reduced_system_kernels_and_modes_agree" > $O/r2c_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2c_quick.log
tail -30 $O/r2c_quick.log
timeout 600 python -m pytest tests/test_reference_golden.py -x -q -m gpu -k "cfg3_full_size_members" > $O/r2c_golden.log 2>&1
echo "golden rc=$?" | tee -a $O/r2c_golden.log
tail -15 $O/r2c_golden.log
timeout 600 python profiles/rows_probe.py 4096 "" "duo=0" "duo_wpc=2" "duo_wpc=1" > $O/r2c_probe.txt 2>&1
timeout 300 python profiles/rows_probe.py 8192 "" "duo=0" >> $O/r2c_probe.txt 2>&1
timeout 300 python profiles/rows_probe.py 512 "" "duo=0" >> $O/r2c_probe.txt 2>&1
cat $O/r2c_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2c_launches.csv python profiles/rows_probe.py 4096 "" > $O/r2c_ncu_list.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2c_launches.csv')) if len(r) > 5]
for r in rows[-12:]:
    print(r[4][:60], r[-1])
PY
