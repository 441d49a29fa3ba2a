import os, sys, time
ROOT=os.getcwd()
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")): sys.path.insert(0, p)
import numpy as np, torch
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL as TQ
from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL as Chain
gam = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]
tq = TQ(10, 500, 15, 50e6, 648e-9 / 530, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6, 2 * np.pi * 100e6, gam)
ch = Chain(16, 4000, 1e-9, 10, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
for name, prop, shape in (("cfg2", tq, (10,5)), ("cfg4", ch, (16,2))):
    rng = np.random.default_rng(0)
    a, b, c = rng.uniform(-1, 1, shape), rng.uniform(-1, 1, shape), rng.uniform(0, 0.5, shape)
    abc = prop._pack(a, b, c); out = torch.empty(3 + abc.numel(), dtype=torch.float64, device=abc.device)
    ts=[]
    for i in range(8):
        torch.cuda.synchronize(); t0=time.perf_counter()
        prop.cost_and_grad_packed(abc, out)
        torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    print(name, "wall ms per call:", " ".join("%.2f"%t for t in ts))
