#!/usr/bin/env bash
# Round 2, GPU call 2: variants of the rows kernel + ncu of k_chain_rows (set full + source), GPU tests again.
O=gpurun_out
mkdir -p $O
timeout 600 python profiles/rows_probe.py 4096 "" "rows_tma=0" "store=0" "spb=5" "spb=15" "rows=0" > $O/r2b_probe.txt 2>&1
timeout 300 python profiles/rows_probe.py 512 "" "store=0" "rows=0" >> $O/r2b_probe.txt 2>&1
cat $O/r2b_probe.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_rows -c 1 -o $O/r2b_rows python profiles/rows_probe.py 4096 "" > $O/r2b_ncu.log 2>&1
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x > $O/r2b_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2b_tests.log
tail -5 $O/r2b_tests.log
