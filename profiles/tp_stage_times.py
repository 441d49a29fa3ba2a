"""Where the time of a single-sample evaluation goes (cfg 2: notebook-02 CZ grid): the C-ABI call sequence is replayed
with CUDA events between the stages by timing cost-only (column sweeps + combine) against cost+gradient, and the
pulse map alone.  usage: python profiles/tp_stage_times.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL as TQ  # noqa: E402

gam = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]
rng = np.random.default_rng(0)
a, b, c = rng.uniform(-1, 1, (10, 5)), rng.uniform(-1, 1, (10, 5)), rng.uniform(0, 0.5, (10, 5))
tq = TQ(10, 500, 15, 50e6, 648e-9 / 530, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6,
        2 * np.pi * 100e6, gam)
abc = tq._pack(a, b, c)
plan = tq._plan


def timed(fn, reps=200):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


wave = plan.physical_amplitudes(abc)
print("chain info", plan.chain_info() if plan.launches else None)
print(f"pulse map forward            {timed(lambda: plan.physical_amplitudes(abc)):8.1f} us")
print(f"cost only  (map + A + B)     {timed(lambda: plan.cost(abc)):8.1f} us")
print(f"cost+grad  (map + A + B + C + vjp) {timed(lambda: plan.cost_grad(abc)):8.1f} us")
print(f"from wave: cost only         {timed(lambda: plan.cost_grad_from_wave(wave, want_grad=False)):8.1f} us")
print(f"from wave: cost+grad         {timed(lambda: plan.cost_grad_from_wave(wave, want_grad=True)):8.1f} us")
print("chain info", plan.chain_info())
