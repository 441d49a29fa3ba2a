#!/usr/bin/env bash
# Round 2, run 22: ncu of the samples-minor kernel after the list-chunk change, in both precisions (cfg-5 seed of 128
# samples = 16 CTAs), and the launch list of one cfg-5 / cfg-4 evaluation (k_wide_degrees, k_wide_phase durations).
O=gpurun_out
mkdir -p $O
cat > /tmp/c64_seed.py <<'PY'
import os, sys
ROOT = os.getcwd()
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL as Chain
prec = sys.argv[1]
rng = np.random.default_rng(100)
prop = Chain(16, 256, 2e-6 / 256, 12, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
nrng = np.random.default_rng(1234)
prop.set_ensemble(del_total=nrng.normal(0, 2 * np.pi * 0.1e6, 128), rabi_scale=1 + nrng.normal(0, 0.01, 128))
prop.set_precision(prec)
a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
for _ in range(3):
    out = prop.cost_and_grad_packed(prop._pack(a, b, c))
torch.cuda.synchronize()
print(prec, float(out[0]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_wide -s 2 -c 1 -o $O/r2i_wide_c128 python /tmp/c64_seed.py complex128 > $O/r2i_ncu_wide_c128.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_wide -s 2 -c 1 -o $O/r2i_wide_c64 python /tmp/c64_seed.py complex64 > $O/r2i_ncu_wide_c64.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2i_launches_cfg5.csv python /tmp/c64_seed.py complex128 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2i_launches_cfg4.csv python profiles/cfg5_probe.py big4 1 > /dev/null 2>&1
ls -la $O/r2i_*
