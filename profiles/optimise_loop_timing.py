"""Wall-clock of the notebook-style optimisation loop (notebook 02 cell 13: 1000 GD steps on the CZ notebook grid),
Python loop with a host sync per step vs prop.optimize (CUDA-graph replay).  python profiles/optimise_loop_timing.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL, single_optimization_step

gam = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]
rng = np.random.default_rng(0)
a, b, c = rng.uniform(-1, 1, (10, 5)), rng.uniform(-1, 1, (10, 5)), rng.uniform(0, 0.5, (10, 5))
prop = PropagatorVL(10, 500, 15, 50e6, 648e-9 / 530, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6,
                    2 * np.pi * 100e6, gam)
steps, lr = 1000, 0.02
ta, tb, tc = a, b, c
single_optimization_step(prop, ta, tb, tc, lr)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(steps):
    cost, ta, tb, tc = single_optimization_step(prop, ta, tb, tc, lr)
    float(cost)                                     # the notebook's per-step host read
t_py = time.perf_counter() - t0
prop.optimize(a, b, c, 5, lr)
torch.cuda.synchronize(); t0 = time.perf_counter()
hist, a2, b2, c2 = prop.optimize(a, b, c, steps, lr)
torch.cuda.synchronize(); t_graph = time.perf_counter() - t0
print(f"1000 GD steps, CZ notebook grid: python loop {t_py:.3f} s ({1e3*t_py/steps:.3f} ms/step), "
      f"graph replay {t_graph:.3f} s ({1e3*t_graph/steps:.3f} ms/step); cost {float(hist[0]):.6f} -> {float(hist[-1]):.6f}; "
      f"final costs agree to {abs(float(hist[-1]) - float(cost)):.2e}")
