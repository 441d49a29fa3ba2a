import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/optimal-control-rydberg_b200')
import numpy as np, torch
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL as TQ
gam = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]
rng = np.random.default_rng(0)
a, b, c = rng.uniform(-1, 1, (10, 5)), rng.uniform(-1, 1, (10, 5)), rng.uniform(0, 0.5, (10, 5))
tq = TQ(10, 500, 15, 50e6, 648e-9 / 530, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6, 2 * np.pi * 100e6, gam)
abc = tq._pack(a, b, c)
for _ in range(3): out = tq.cost_and_grad_packed(abc)
torch.cuda.synchronize()
print(float(out[0]))
