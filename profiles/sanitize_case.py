"""Smoke-size problems for compute-sanitizer (memcheck / racecheck / synccheck), one kernel family per case:

    compute-sanitizer --tool racecheck python profiles/sanitize_case.py duo

cases: duo, duo_wpc1 (chain_duo.cu), aug, aug_recompute, aug_c64 (chain_aug.cu),
tp (chain_small.cu time-parallel single sample), wide, wide_tp (chain_wide.cu), lindblad.
Each case prints the cost it computed so a sanitizer run that silently did nothing is visible."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import qocvl  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "duo"
GAM = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]


def cz(n_samples, nt=40, **opts):
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    pad = int(0.03 * nt)
    with qocvl.options(**opts):
        prop = PropagatorVL(4, nt, pad, 50e6, 648e-9 / 530, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6),
                            2 * np.pi * 100e6, 2 * np.pi * 100e6, GAM)
    rng = np.random.default_rng(0)
    a, b, c = rng.uniform(-1, 1, (4, 5)), rng.uniform(-1, 1, (4, 5)), rng.uniform(0, 0.5, (4, 5))
    if n_samples > 1:
        nrng = np.random.default_rng(1)
        prop.set_ensemble(del_total=nrng.normal(0, 6e5, n_samples), rabi_scale=1 + nrng.normal(0, 0.01, n_samples))
    return prop, a, b, c


def chain(n_samples, n_atoms=6, nt=24, **opts):
    from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL
    with qocvl.options(**opts):
        prop = PropagatorVL(4, nt, 20e-9, n_atoms, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    rng = np.random.default_rng(0)
    a, b, c = rng.uniform(-1, 1, (4, 2)), rng.uniform(-1, 1, (4, 2)), rng.uniform(0, 0.5, (4, 2))
    if n_samples > 1:
        nrng = np.random.default_rng(1)
        prop.set_ensemble(del_total=nrng.normal(0, 6e5, n_samples), rabi_scale=1 + nrng.normal(0, 0.01, n_samples))
    return prop, a, b, c


if case == "lindblad":
    from quantum_optimal_control.two_qubit.qutip_five_level import Mesolve_5lvl_t  # noqa: F401
    from qocvl import lindblad as ql
    import inspect
    print("lindblad module:", [n for n, _ in inspect.getmembers(ql, inspect.isfunction)][:5])
    sys.exit(0)
builders = {
    "duo": lambda: cz(7), "duo_wpc1": lambda: cz(7, duo_wpc=1),
    "aug": lambda: cz(7, duo=0), "aug_recompute": lambda: cz(7, duo=0, store=0), "aug_c64": lambda: cz(7),
    "tp": lambda: cz(1, nt=64), "wide": lambda: chain(9), "wide_tp": lambda: chain(1, nt=80),
}
prop, a, b, c = builders[case]()
if case == "aug_c64":
    prop.set_precision("complex64")
out = prop.cost_and_grad_packed(prop._pack(a, b, c)).cpu().numpy()
prop._plan.check()
print(case, prop._plan.chain_info()["path"], "cost", out[0], "grad max", np.abs(out[3:]).max())
