"""cfg 1 / cfg 2 (the reference's own notebooks: ONE sample) on the time-parallel path of chain_small.cu against the
time-parallel form of the two-lane ensemble kernels (chain_duo.cu) with explicit chunk counts:

    python profiles/single_sample_probe.py "" "tp=0,strong_chunks=16" "tp=0,strong_chunks=32" "tp=0,strong_chunks=64"
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import qocvl  # noqa: E402
from measure_configs import timed  # noqa: E402
from quantum_optimal_control.state_transfer.propagator_vl import PropagatorVL as ST  # noqa: E402
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL as TQ  # noqa: E402

gam = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]
for spec in (sys.argv[1:] or [""]):
    kv = {k: int(v) for k, v in (item.split("=") for item in spec.split(",") if item)}
    rng = np.random.RandomState(1)
    a = rng.uniform(-1, 1, (10, 5))
    b = np.concatenate([rng.uniform(-1, 1, (10, 1)), rng.uniform(0, 1, (10, 2)), rng.uniform(-1, 0, (10, 2))], axis=1)
    c = rng.uniform(0, 0.1, (10, 5))
    with qocvl.options(**kv):
        st = ST(10, 500, 4e-10, 0, 0, 2 * np.pi * 127e6, 2 * np.pi * 127e6, 1 / 150e-6, 1 / 118e-9)
        tq = TQ(10, 500, 15, 50e6, 648e-9 / 530, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6,
                2 * np.pi * 100e6, gam)
    ms1, k1, _, c1 = timed(st, a, b, c)
    r1 = timed.last_repeat_ms
    rng = np.random.default_rng(0)
    a2, b2, c2 = rng.uniform(-1, 1, (10, 5)), rng.uniform(-1, 1, (10, 5)), rng.uniform(0, 0.5, (10, 5))
    ms2, k2, _, c2v = timed(tq, a2, b2, c2)
    print(f"opts[{spec}] repeated calls with the same buffers (cached graph): cfg1 {r1:.3f} ms | cfg2 {timed.last_repeat_ms:.3f} ms")
    print(f"opts[{spec}] cfg1 {ms1:.3f} ms (chain {k1:.3f}, path {st._plan.chain_info()['path']}, chunks {st._plan.slice_chunks}, "
          f"cost {c1:.12f}) | cfg2 {ms2:.3f} ms (chain {k2:.3f}, path {tq._plan.chain_info()['path']}, chunks "
          f"{tq._plan.slice_chunks}, cost {c2v:.12f})", flush=True)
