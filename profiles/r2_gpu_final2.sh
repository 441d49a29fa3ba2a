#!/usr/bin/env bash
# Round 2, final measurement call of the build with per-role Taylor degrees / phase shifts / complex64 wide kernel:
# bench lines (both arms), launch list, ncu --set full of the two sweeps and of the samples-minor kernel, strong-scaling
# shard timings, wide-kernel configs.  Everything lands in gpurun_out/r2f_*; summaries are copied into profiles/ by hand.
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_n1_reference_arm.json 2> $O/r2f_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-c64 > $O/r2f_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo_reverse -c 1 -o $O/r2f_rev python profiles/ensemble_probe.py 4096 "" > $O/r2f_ncu_rev.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo_forward -c 1 -o $O/r2f_fwd python profiles/ensemble_probe.py 4096 "" > $O/r2f_ncu_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_wide -c 1 -o $O/r2f_wide python profiles/cfg5_probe.py cfg5 128 > $O/r2f_ncu_wide.log 2>&1
( for S in 4096 2048 1024 512; do timeout 300 python profiles/ensemble_probe.py $S "" "strong_chunks=1"; done
  timeout 300 python profiles/ensemble_probe.py 4096 "duo_roleq=2" "duo_roleq=1" "duo_roleq=0" "duo=0" "duo_nolean=1" ) > $O/r2f_strong.txt 2>&1
( timeout 300 python profiles/cfg5_probe.py cfg5 512 c64; timeout 300 python profiles/cfg5_probe.py big4 1 c64
  timeout 600 python profiles/cfg5_multistart.py ) > $O/r2f_wide.txt 2>&1
timeout 300 python profiles/single_sample_probe.py > $O/r2f_single.txt 2>&1
tail -3 $O/r2f_strong.txt; tail -4 $O/r2f_wide.txt; tail -c 300 $O/r2f_bench_n1.err
