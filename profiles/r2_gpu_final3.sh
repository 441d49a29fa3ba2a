#!/usr/bin/env bash
# Round 2, last measurement call (the kernels are those of r2_gpu_final2.sh; new since then: cached evaluation graphs,
# 37-seed multistart sample, repeated-call timings in `configs`): the driver's own sequence -- GPU tests, smoke, bench
# (both arms), launch list.
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/r2g_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -4 | tee $O/r2g_smoke.log
timeout 900 python bench.py > $O/r2g_bench_n1.json 2> $O/r2g_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2g_bench_n1_reference_arm.json 2> $O/r2g_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-c64 > $O/r2g_ncu_list.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'])
for k,v in d['configs'].items(): print(k, {kk:v[kk] for kk in v if kk.startswith('ms_') or kk in ('gpu_launches_per_eval','concurrent_replicas','full_config_seconds_1gpu')})
PY
tail -c 300 $O/r2g_bench_n1.err
