#!/usr/bin/env bash
# Round 2, GPU call 8: Xbar sums through shared memory (default build); halo through shared memory (variant hs)
O=gpurun_out
mkdir -p $O
V=optimal-control-rydberg_b200/build/variants
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_reference_golden.py -x -q -m gpu -k "reduced_system_kernels or cfg3_full_size_members" > $O/r2j_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2j_quick.log
QOCVL_LIB=$V/libqocvl_hs.so timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "reduced_system_kernels" > $O/r2j_quick_hs.log 2>&1
echo "quick hs rc=$?" | tee -a $O/r2j_quick_hs.log
for v in base hs; do
  lib=optimal-control-rydberg_b200/libqocvl.so; [ $v = base ] || lib=$V/libqocvl_$v.so
  QOCVL_LIB=$lib timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2j_launches_$v.csv python profiles/ensemble_probe.py 4096 "" > /dev/null 2>&1
  grep -E "k_duo_(forward|reverse)" $O/r2j_launches_$v.csv | tail -2 | awk -F'","' -v v=$v '{print v, $5, $NF}'
  QOCVL_LIB=$lib timeout 300 python profiles/ensemble_probe.py 4096 "" | sed "s/^/$v /"
done 2>&1 | tee $O/r2j_probe.txt
