#!/usr/bin/env bash
# Round 2, final measurement call: bench lines (both arms), launch list, ncu --set full of the two sweeps, strong-scaling
# shard timings, microbenchmarks.  Everything lands in gpurun_out/r2z_*; the summaries are copied into profiles/ by hand.
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/r2z_bench_n1.json 2> $O/r2z_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2z_bench_n1_reference_arm.json 2> $O/r2z_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-c64 > $O/r2z_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo_reverse -c 1 -o $O/r2z_rev python profiles/ensemble_probe.py 4096 "" > $O/r2z_ncu_rev.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_duo_forward -c 1 -o $O/r2z_fwd python profiles/ensemble_probe.py 4096 "" > $O/r2z_ncu_fwd.log 2>&1
( for S in 4096 2048 1024 512; do timeout 300 python profiles/ensemble_probe.py $S "" "strong_chunks=1"; done
  timeout 300 python profiles/ensemble_probe.py 4096 "duo=0" "duo=0,store=0" "duo_nolean=1" ) > $O/r2z_strong.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2z_launches_512.csv python profiles/ensemble_probe.py 512 "" > /dev/null 2>&1
( cd profiles/micro && ./dfma_operands && ./latency ) > $O/r2z_micro.txt 2>&1
tail -3 $O/r2z_strong.txt; tail -c 300 $O/r2z_bench_n1.err
