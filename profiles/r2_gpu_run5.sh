#!/usr/bin/env bash
# Round 2, GPU call 5: split accumulation chains (default build) and the unroll-by-2 variant.
O=gpurun_out
mkdir -p $O
V=optimal-control-rydberg_b200/build/variants
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "reduced_system_kernels" > $O/r2e_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2e_quick.log
( timeout 300 python profiles/ensemble_probe.py 4096 "" ; for v in u2; do QOCVL_LIB=$V/libqocvl_$v.so timeout 300 python profiles/ensemble_probe.py 4096 "" | sed "s/^/$v /"; done ) > $O/r2e_probe.txt 2>&1
cat $O/r2e_probe.txt
for v in base u2; do
  lib=optimal-control-rydberg_b200/libqocvl.so; [ $v = base ] || lib=$V/libqocvl_$v.so
  QOCVL_LIB=$lib timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2e_launches_$v.csv python profiles/ensemble_probe.py 4096 "" > /dev/null 2>&1
  grep -E "k_duo_(forward|reverse)" $O/r2e_launches_$v.csv | tail -2 | awk -F'","' -v v=$v '{print v, $5, $NF}'
done
