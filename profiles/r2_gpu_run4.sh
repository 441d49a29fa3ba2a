#!/usr/bin/env bash
# Round 2, GPU call 4: chain_duo.cu -- agreement / golden tests, then ncu --set full of both sweeps.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_reference_golden.py tests/test_gpu_fullsize.py -x -q -m gpu -k "reduced or cfg3 or sharding or ensemble" > $O/r2d_quick.log 2>&1
echo "quick rc=$?" | tee -a $O/r2d_quick.log
tail -8 $O/r2d_quick.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo_reverse -c 1 -o $O/r2d_rev python profiles/ensemble_probe.py 4096 "" > $O/r2d_ncu_rev.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo_forward -c 1 -o $O/r2d_fwd python profiles/ensemble_probe.py 4096 "" > $O/r2d_ncu_fwd.log 2>&1
tail -3 $O/r2d_ncu_rev.log $O/r2d_ncu_fwd.log
