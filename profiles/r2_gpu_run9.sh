#!/usr/bin/env bash
# Round 2, GPU call 9: ncu --set full of k_duo_reverse / k_duo_forward (current default build)
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo_reverse -c 1 -o $O/r2q_rev python profiles/ensemble_probe.py 4096 "" > $O/r2q_ncu_rev.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo_forward -c 1 -o $O/r2q_fwd python profiles/ensemble_probe.py 4096 "" > $O/r2q_ncu_fwd.log 2>&1
ls -la $O/r2q_*
