"""cfg-3 per-step time of the ensemble kernels under explicit plan options:

    python profiles/ensemble_probe.py 4096 "" "duo_wpc=1" "duo=0" "duo=0,store=0"

Each argument after the sample count is one comma-separated option set (qocvl_set_option names); prints ms per
cost+gradient (CUDA events, 2 warm-up + 3 timed) and the chain-kernel time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import qocvl  # noqa: E402
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sets = sys.argv[2:] or [""]
NG, NT = 16, 1888
pad = int(0.03 * NT)
M = NT + 2 * pad
rng = np.random.default_rng(0)
a, b, c = rng.uniform(-1, 1, (NG, 5)), rng.uniform(-1, 1, (NG, 5)), rng.uniform(0, 0.5, (NG, 5))
nrng = np.random.default_rng(1234)
dl, eps = nrng.normal(0.0, 2 * np.pi * 0.1e6, max(S, 4096))[:S], nrng.normal(0.0, 0.01, max(S, 4096))[:S]
for spec in sets:
    kv = {k: int(v) for k, v in (item.split("=") for item in spec.split(",") if item)}
    with qocvl.options(**kv):
        prop = PropagatorVL(NG, NT, pad, 50e6, 648e-9 / M, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6,
                            2 * np.pi * 100e6, [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6])
    prop.set_ensemble(del_total=dl, rabi_scale=1.0 + eps)
    abc = prop._pack(a, b, c)
    prop._plan.set_timing(True)
    for _ in range(2):
        out = prop.cost_and_grad_packed(abc)
    prop._plan.check()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = prop.cost_and_grad_packed(abc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    info = prop._plan.chain_info()
    print(f"S={S} opts[{spec}] {ms:.3f} ms/step  chain {prop._plan.last_chain_ms():.3f} ms  {M * S / ms * 1e3:.3e} units/s  "
          f"path {info['path']} threads {info['threads']} ctas {info['ctas']} cost {float(out[0]):.12f} "
          f"fwd {prop._plan.last_stage_ms(0):.3f} rev {prop._plan.last_stage_ms(1):.3f} ms  degrees {prop._plan.role_degree_means()}", flush=True)
    del prop
