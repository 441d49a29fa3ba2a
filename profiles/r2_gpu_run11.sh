#!/usr/bin/env bash
# Round 2, GPU call 11: time-parallel evaluation of small ensembles (slice-axis chunks) on the two-lane kernels
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_reference_golden.py tests/test_gpu_fullsize.py -x -q -m gpu -k "reduced_system_kernels or cfg3_full_size_members or sharding" > $O/r2m_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2m_quick.log
tail -12 $O/r2m_quick.log
for S in 512 1024 2048; do
  timeout 300 python profiles/ensemble_probe.py $S "" "strong_chunks=1" "strong_chunks=2" "strong_chunks=4" "strong_chunks=8" "strong_chunks=16" "duo=0"
done 2>&1 | tee $O/r2m_probe.txt
