"""cfg 3 (robust CZ, 2000 slices, N = 16) cost+gradient in complex128 or complex64, for ncu captures and for scans of the
launch geometry: `cfg3_probe.py [c128|c64] [samples] [QOCVL_SPB values ...]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "c128"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spbs = [int(x) for x in sys.argv[3:]] or [0]
NT, NG = 1888, 16
pad = int(0.03 * NT)
M = NT + 2 * pad
prop = PropagatorVL(NG, NT, pad, 50e6, 648e-9 / M, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6,
                    2 * np.pi * 100e6, [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6])
prop.set_precision("complex64" if dtype == "c64" else "complex128")
rng = np.random.default_rng(0)
a, b, c = rng.uniform(-1, 1, (NG, 5)), rng.uniform(-1, 1, (NG, 5)), rng.uniform(0, 0.5, (NG, 5))
nrng = np.random.default_rng(1234)
dl, eps = nrng.normal(0.0, 2 * np.pi * 0.1e6, S), nrng.normal(0.0, 0.01, S)
abc = prop._pack(a, b, c)
for spb in spbs:
    if spb:
        prop._plan.set_option("spb", spb)
    prop.set_ensemble(del_total=dl, rabi_scale=1.0 + eps)      # re-configures the launch
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
    print(f"{dtype} S={S} spb={spb or 'auto'}: {ms:.3f} ms ({M * S / ms * 1e3:.3e} units/s), chain kernels "
          f"{prop._plan.last_chain_ms():.3f} ms, cost {float(out[0]):.9f}, cap {prop._plan.taylor_cap}, "
          f"{info['threads']} thr x {info['ctas']} CTAs")
