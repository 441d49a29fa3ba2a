"""Wall time of the GPU master-equation validation at the notebook-05 size (530 time points over 648 ns, two 5-level
atoms, 8 collapse operators, default sub-stepping), and its agreement with the CPU ODE oracle."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from oracle import lindblad as ol  # noqa: E402
from quantum_optimal_control.two_qubit.qutip_five_level import Mesolve_5lvl_t  # noqa: E402

M, T = 530, 648e-9
tl = np.linspace(0.0, T, M)
s = tl / T
pulses = [np.ones(M), np.sin(0.5 * np.pi * s), np.zeros(M), np.abs(np.cos(0.5 * np.pi * s)), np.zeros(M)]   # NB05 cell 22
g0, g1 = ol.ket(ol.G0), ol.ket(ol.G1)
psi0 = 0.5 * (np.kron(g0, g0) + np.kron(g0, g1) + np.kron(g1, g0) + np.kron(g1, g1))
psit = 0.5 * (np.kron(g0, g0) + np.kron(g0, g1) + np.kron(g1, g0) - np.kron(g1, g1))
gam = [0.1, 1 / 118e-9, 1 / 1.5e-3, 1 / 170e-6]
V, R, D = 2 * np.pi * 10e6, 2 * np.pi * 100e6, 2 * np.pi * (-35.7e6)
args = [psi0, psit, V, R, R, pulses[1], pulses[3], pulses[2], pulses[4], pulses[0], D, 0.0, gam]
Mesolve_5lvl_t(tl, args, output_states=True)
t0 = time.perf_counter()
res = Mesolve_5lvl_t(tl, args, output_states=True)
t1 = time.perf_counter()
print(f"GPU Mesolve_5lvl_t: {1e3 * (t1 - t0):.1f} ms wall (host setup + kernel + copies), fidelity {res.expect[1][-1]:.6f}")
t0 = time.perf_counter()
ref = ol.evolve_ode(tl, pulses, ol.hamiltonian_pieces(V, R, R, D, 0.0), ol.collapse_ops(gam), np.outer(psi0, psi0.conj()),
                    rtol=1e-8, atol=1e-10)
t1 = time.perf_counter()
print(f"CPU solve_ivp oracle: {t1 - t0:.1f} s, max |rho_gpu - rho_ode| = {np.abs(res.final_state - ref[-1]).max():.2e}")
