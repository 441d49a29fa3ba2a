"""Master-equation validation path: the GPU Lindblad propagator behind the reference's
``Mesolve_5lvl_t`` against the CPU oracle (oracle/lindblad.py), and the oracle's two integrators against each other."""
import numpy as np
import pytest

from oracle import lindblad as ol

GAMMAS = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]
V_INT, RABI, DELTA_I = 2 * np.pi * 10e6, 2 * np.pi * 100e6, 2 * np.pi * (-35.7e6)


def setup(M=25, T=120e-9, gammas=GAMMAS):
    time = np.linspace(0.0, T, M)
    s = time / T
    pulses = [-np.ones(M) * (0.6 + 0.4 * s),                       # Delta_i pulse
              np.sin(0.5 * np.pi * s) * np.cos(0.3 * s), np.sin(0.5 * np.pi * s) * np.sin(0.3 * s),
              np.abs(np.cos(0.5 * np.pi * s)) * np.cos(0.2 * s), -np.abs(np.cos(0.5 * np.pi * s)) * np.sin(0.2 * s)]
    g0, g1 = ol.ket(ol.G0), ol.ket(ol.G1)
    psi0 = 0.5 * (np.kron(g0, g0) + np.kron(g0, g1) + np.kron(g1, g0) + np.kron(g1, g1))
    psit = 0.5 * (np.kron(g0, g0) + np.kron(g0, g1) + np.kron(g1, g0) - np.kron(g1, g1))
    return time, pulses, psi0, psit, gammas


def reference_args(time, pulses, psi0, psit, gammas):
    d, ire, iim, rre, rim = pulses
    return [psi0, psit, V_INT, RABI, RABI, ire, rre, iim, rim, d, DELTA_I, 0.0, gammas]


def test_oracle_midpoint_scheme_converges_to_the_ode_solution():
    time, pulses, psi0, psit, gammas = setup(M=13, T=40e-9)
    Hs, cs = ol.hamiltonian_pieces(V_INT, RABI, RABI, DELTA_I, 0.0), ol.collapse_ops(gammas)
    rho0 = np.outer(psi0, psi0.conj())
    exact = ol.evolve_ode(time, pulses, Hs, cs, rho0)
    errs = [np.abs(ol.evolve_midpoint(time, pulses, Hs, cs, rho0, sub) - exact).max() for sub in (4, 8, 16)]
    assert errs[2] < 5e-5 and errs[0] / errs[1] > 3.0 and errs[1] / errs[2] > 3.0     # second order
    assert abs(np.trace(exact[-1]).real - 1.0) < 1e-8


def test_mirror_builds_the_same_operators_as_the_oracle():
    from quantum_optimal_control.two_qubit import qutip_five_level as q5, qutip_simulation as qs
    time, pulses, psi0, psit, gammas = setup()
    H = q5.QTHamiltonian(reference_args(time, pulses, psi0, psit, gammas))
    Hs = ol.hamiltonian_pieces(V_INT, RABI, RABI, DELTA_I, 0.0)
    for mine, ref in zip([H[0]] + [h for h, _ in H[1:]], Hs):
        assert np.array_equal(mine, ref)
    for mine, ref in zip(q5.collapse_operators(gammas), ol.collapse_ops(gammas)):
        assert np.array_equal(mine, ref)
    for mine, ref in zip(q5.liouvillian_pieces(H, q5.collapse_operators(gammas)), ol.superoperators(Hs, ol.collapse_ops(gammas))):
        assert np.abs(mine - ref).max() < 1e-6 * np.abs(ref).max()
    assert qs.Mesolve_5lvl_t is q5.Mesolve_5lvl_t and len(q5.gen_n_level_atom_basis(5)) == 5


@pytest.mark.gpu
def test_gpu_master_equation_matches_oracle():
    from quantum_optimal_control.two_qubit.qutip_five_level import Mesolve_5lvl_t
    time, pulses, psi0, psit, gammas = setup()
    args = reference_args(time, pulses, psi0, psit, gammas)
    res = Mesolve_5lvl_t(time, args, output_states=True, substeps=24)
    Hs, cs = ol.hamiltonian_pieces(V_INT, RABI, RABI, DELTA_I, 0.0), ol.collapse_ops(gammas)
    rho0 = np.outer(psi0, psi0.conj())
    ref = ol.evolve_midpoint(time, pulses, Hs, cs, rho0, 24)           # same discretisation, SciPy exponential action
    assert np.abs(res.final_state - ref[-1]).max() < 1e-11
    pops = [np.einsum("i,tij,j->t", v.conj(), ref, v).real for v in (psi0, psit)]
    assert np.abs(res.expect[0] - pops[0]).max() < 1e-11 and np.abs(res.expect[1] - pops[1]).max() < 1e-11
    assert len(res.expect) == 4 and res.expect[0].shape == time.shape and abs(res.expect[0][0] - 1.0) < 1e-14
    exact = ol.evolve_ode(time, pulses, Hs, cs, rho0)                  # independent integrator
    auto = Mesolve_5lvl_t(time, args, output_states=True)              # default sub-stepping
    assert np.abs(auto.final_state - exact[-1]).max() < 2e-5
    assert abs(np.trace(auto.final_state).real - 1.0) < 1e-10
    assert np.abs(auto.final_state - auto.final_state.conj().T).max() < 1e-12


@pytest.mark.gpu
def test_gpu_master_equation_without_decay_equals_the_unitary_gate_fidelity():
    """Gamma = 0: the dark level is never populated and <psi_t| rho(T) |psi_t> must equal |<psi_t| U |psi_0>|^2
    with U from the 4-level engine's own propagate() (TQ:489-535) on the same pulses."""
    from oracle import two_qubit as otq
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    from quantum_optimal_control.two_qubit.qutip_five_level import Mesolve_5lvl_t
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    zero = [0.0, 0.0, 0.0, 0.0]
    prop = PropagatorVL(6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val,
                        po.Rabi_r_val, zero)
    wave = prop.return_physical_amplitudes(a, b, c).cpu().numpy()            # [M, 5]
    U = prop.propagate(a, b, c).cpu().numpy()[:16, :16]
    M = wave.shape[0]
    # the engine holds each slice's amplitudes constant over [k dt, (k+1) dt): give the master equation the same
    # piecewise-constant pulses as functions of time (QuTiP's function-format coefficients)
    dt = po.delta_t
    step = lambda arr: (lambda t, args: float(arr[min(int(t / dt), M - 1)]))   # noqa: E731
    time = np.linspace(0.0, M * dt, M + 1)
    ire, iim = wave[:, 1] * np.cos(np.pi * wave[:, 2]), wave[:, 1] * np.sin(np.pi * wave[:, 2])
    rre, rim = wave[:, 3] * np.cos(np.pi * wave[:, 4]), wave[:, 3] * np.sin(np.pi * wave[:, 4])
    g0, g1 = ol.ket(ol.G0), ol.ket(ol.G1)
    psi0 = 0.5 * (np.kron(g0, g0) + np.kron(g0, g1) + np.kron(g1, g0) + np.kron(g1, g1))
    psit = 0.5 * (np.kron(g0, g0) + np.kron(g0, g1) + np.kron(g1, g0) - np.kron(g1, g1))
    args = [psi0, psit, po.Vint, po.Rabi_i_val, po.Rabi_r_val, step(ire), step(rre), step(iim), step(rim),
            step(wave[:, 0]), po.Delta_i_val, 0.0, zero]
    res = Mesolve_5lvl_t(time, args, output_states=True, substeps=8)
    four = lambda v: v.reshape(5, 5)[:4, :4].reshape(-1)                      # noqa: E731
    fid_unitary = abs(np.vdot(four(psit), U @ four(psi0))) ** 2
    assert abs(res.expect[1][-1] - fid_unitary) < 1e-9
