#!/usr/bin/env bash
# Round 2, run 19: bench line + launch list of the build with per-role Taylor degrees and the complex64 wide kernel.
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/r2ad_bench_n1.json 2> $O/r2ad_bench_n1.err; echo "bench rc=$?"
tail -c 400 $O/r2ad_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ad_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['value'], d['ms_per_step'], d['e2e']['value'], 'frac', r['frac'], r['frac_of_register_operand_peak'], 'rev ms', r['kernel_ms'], r['taylor_degrees']['per_role'], 'eval frac', r['evaluation']['frac_of_peak'], r['evaluation']['m2_frac_of_peak_every_entry'])
print(d['parity']['ok'], d['complex64']['ms_per_step'])
for k,v in d['configs'].items(): print(k, {kk:v[kk] for kk in v if kk in ('ms_per_eval','ms_per_seed','kernel_path')})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2ad_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-c64 > $O/r2ad_ncu_list.log 2>&1
