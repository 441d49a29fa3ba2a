"""exponentials() / propagate() of the blockade-chain model at the BASELINE sizes (cfg 4: n = 144, M = 4000; cfg 5: n = 377,
M = 256) on the block-triangular DMMA path of dense_blocks.cu: wall time (CUDA events around the C-ABI call) and the
rate of the batched complex GEMM kernel.   python profiles/dense_blocks_timing.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL as Chain  # noqa: E402

for name, n_atoms, nt, dt in (("cfg4 n=144 M=4000", 10, 4000, 1e-9), ("cfg5 n=377 M=256", 12, 256, 2e-6 / 256)):
    rng = np.random.default_rng(7)
    a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
    prop = Chain(16, nt, dt, n_atoms, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    n = prop.dim
    for what in ("propagate", "exponentials"):
        fn = getattr(prop, what)
        out = fn(a, b, c)
        torch.cuda.synchronize()
        l0 = prop._plan.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(a, b, c)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        gemms = prop._plan.launches - l0
        print(f"{name} {what}: {ms:.1f} ms, {gemms} launches, output {tuple(out.shape)}", flush=True)
        del out
    vl = prop.propagate(a, b, c)
    o = prop.cost_and_grad_packed(prop._pack(a, b, c)).cpu().numpy()
    print(f"   |W| max {float(vl[:n, n:].abs().max()):.3e}; chain cost {o[0]:.12f}")
