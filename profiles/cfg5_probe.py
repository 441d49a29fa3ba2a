"""One cost+gradient of one cfg-5 seed (12-atom blockade chain, n = 377, M = 256, 512 noise samples) and, with
`big4`, of cfg 4 (10 atoms, n = 144, M = 4000, one sample): the workload of the big-n kernel, for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL as Chain  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 512
if which == "cfg5":
    rng = np.random.default_rng(100)
    prop = Chain(16, 256, 2e-6 / 256, 12, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    nrng = np.random.default_rng(1234)
    prop.set_ensemble(del_total=nrng.normal(0, 2 * np.pi * 0.1e6, n_samples),
                      rabi_scale=1 + nrng.normal(0, 0.01, n_samples))
    units = 256 * n_samples
else:
    rng = np.random.default_rng(7)
    prop = Chain(16, 4000, 1e-9, 10, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    units = 4000
a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
abc = prop._pack(a, b, c)
full = None
for precision in (["complex128", "complex64"] if "c64" in sys.argv[3:] else ["complex128"]):   # 3rd argument "c64": both modes
    prop.set_precision(precision)
    for _ in range(2):
        out = prop.cost_and_grad_packed(abc)
    prop._plan.check()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = prop.cost_and_grad_packed(abc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res = out.detach().cpu().numpy().copy()
    extra = ""
    if full is None:
        full = res
    else:
        extra = (f" | vs complex128: cost {abs(res[0] - full[0]):.2e}, gradient max-rel "
                 f"{np.abs(res[3:] - full[3:]).max() / np.abs(full[3:]).max():.2e}")
    print(which, precision, f"{ms:.2f} ms, {units / ms * 1e3:.3e} units/s, cost {float(out[0]):.6f}, taylor cap {prop._plan.taylor_cap}",
          prop._plan.chain_info(), extra, flush=True)
