"""cfg-5 seed (n = 377, 512 samples) on the samples-minor kernel for every compiled CTA geometry (wide_threads option)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import qocvl  # noqa: E402
from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL as Chain  # noqa: E402

rng = np.random.default_rng(100)
a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
for spec in (sys.argv[1:] or ["", "wide_threads=256", "wide_threads=512", "wide_store=0"]):
    kv = {k: int(v) for k, v in (item.split("=") for item in spec.split(",") if item)}
    with qocvl.options(**kv):
        prop = Chain(16, 256, 2e-6 / 256, 12, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    nrng = np.random.default_rng(1234)
    prop.set_ensemble(del_total=nrng.normal(0, 2 * np.pi * 0.1e6, 512), rabi_scale=1 + nrng.normal(0, 0.01, 512))
    abc = prop._pack(a, b, c)
    try:
        for _ in range(2):
            out = prop.cost_and_grad_packed(abc)
        prop._plan.check()
    except Exception as e:                      # a geometry the problem does not fit
        print(f"opts[{spec}] {type(e).__name__}: {e}")
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        out = prop.cost_and_grad_packed(abc)
    e1.record()
    torch.cuda.synchronize()
    print(f"opts[{spec}] {e0.elapsed_time(e1) / 2:.2f} ms, cost {float(out[0]):.9f}", prop._plan.chain_info(), flush=True)
    del prop
