"""cfg 3: time of the forward sweep alone (qocvl_cost) vs forward + reverse (qocvl_cost_grad)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, NT = 16, 1888
pad = int(0.03 * NT); M = NT + 2 * pad
prop = PropagatorVL(N, NT, pad, 50e6, 648e-9 / M, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6,
                    2 * np.pi * 100e6, [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6])
rng = np.random.default_rng(0)
a, b, c = rng.uniform(-1, 1, (N, 5)), rng.uniform(-1, 1, (N, 5)), rng.uniform(0, 0.5, (N, 5))
nr = np.random.default_rng(1234)
prop.set_ensemble(del_total=nr.normal(0, 2 * np.pi * 0.1e6, S), rabi_scale=1 + nr.normal(0, 0.01, S))
abc = prop._pack(a, b, c)
def t(fn, reps=3):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(f"S={S}: cost only {t(lambda: prop._plan.cost(abc)):.2f} ms, cost+grad {t(lambda: prop._plan.cost_grad(abc)):.2f} ms")
