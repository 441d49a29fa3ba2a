"""GPU replacement of the reference's QuTiP validation modules ``two_qubit/qutip_five_level.py`` and
``two_qubit/qutip_simulation.py`` (same function names and argument lists; no qutip needed).

Two 5-level atoms (g0, g1, i, r, dark), time-dependent Hamiltonian of the CZ pulses, 8 collapse operators
(QF:11-86, QF:88-142 / QS:7-103): the Lindblad master equation is propagated on the GPU by
``qocvl.lindblad.propagate`` (csrc/lindblad.cu) instead of ``qutip.mesolve``.

Time dependence: like QuTiP's array-format coefficients, the pulse samples given on ``time`` are interpolated with a
cubic spline; every interval of ``time`` is cut into ``substeps`` exponential-midpoint steps (second order; the
default sub-stepping keeps |L dt| <= 0.5, i.e. a step error far below QuTiP's default tolerances).  States are plain
NumPy arrays: kets of length 25 (objects with ``.full()`` such as qutip.Qobj are accepted), density matrices 25 x 25.
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import CubicSpline

DIM = 5


def gen_n_level_atom_basis(n: int):
    """QF:4-9 / QS:4-5: [|0>, ..., |n-1>] as NumPy unit vectors."""
    return [np.eye(n, dtype=np.complex128)[k] for k in range(n)]


def tensor(*kets):
    """Kronecker product of kets or operators (stand-in for qutip.tensor in the notebooks)."""
    out = np.asarray(kets[0], dtype=np.complex128)
    for k in kets[1:]:
        out = np.kron(out, np.asarray(k, dtype=np.complex128))
    return out


def _op(row, col):
    m = np.zeros((DIM, DIM), dtype=np.complex128)
    m[row, col] = 1.0
    return m


def _both(single):
    eye = np.eye(DIM, dtype=np.complex128)
    return np.kron(eye, single) + np.kron(single, eye)


def _ket(x):
    x = x.full() if hasattr(x, "full") else x
    return np.asarray(x, dtype=np.complex128).reshape(-1)


def QTHamiltonian(argsME5lvl):
    """QF:11-86: [H_0, [H_1, f_1], ..., [H_5, f_5]] with 25 x 25 NumPy matrices and the pulse sample arrays."""
    (psi0, psi_orth, V_int, Rabi_i_max, Rabi_r_max, Rabi_i_Pulse_Re, Rabi_r_Pulse_Re, Rabi_i_Pulse_Im,
     Rabi_r_Pulse_Im, Delta_i_Pulse, Delta_i_max, del_total, Gammas) = argsME5lvl
    sig_rr, sig_ii, sig_1i, sig_ir = _op(3, 3), _op(2, 2), _op(1, 2), _op(2, 3)
    H0 = _both(-del_total * sig_rr) + V_int * np.kron(sig_rr, sig_rr)
    return [H0,
            [_both(-Delta_i_max * sig_ii), Delta_i_Pulse],
            [_both(-0.5 * Rabi_i_max * (sig_1i + sig_1i.conj().T)), Rabi_i_Pulse_Re],
            [_both(-0.5j * Rabi_i_max * (sig_1i - sig_1i.conj().T)), Rabi_i_Pulse_Im],
            [_both(-0.5 * Rabi_r_max * (sig_ir + sig_ir.conj().T)), Rabi_r_Pulse_Re],
            [_both(-0.5j * Rabi_r_max * (sig_ir - sig_ir.conj().T)), Rabi_r_Pulse_Im]]


def collapse_operators(Gammas):
    """QF:109-118: sqrt(Gamma) * sigma on either atom for sigma in (|0><1|, |1><i|, |i><r|, |d><r|)."""
    eye = np.eye(DIM, dtype=np.complex128)
    ops = []
    for gamma, sig in zip(Gammas, (_op(0, 1), _op(1, 2), _op(2, 3), _op(4, 3))):
        ops.append(np.sqrt(gamma) * np.kron(sig, eye))
        ops.append(np.sqrt(gamma) * np.kron(eye, sig))
    return ops


def liouvillian_pieces(H_list, c_ops):
    """Superoperators on row-major vec(rho):  vec(A rho B) = (A kron B^T) vec(rho).
    -> [L_0 (static Hamiltonian + dissipators), L_1 .. L_5 (one per driven Hamiltonian piece)]."""
    n = H_list[0].shape[0]
    eye = np.eye(n, dtype=np.complex128)
    comm = lambda H: -1j * (np.kron(H, eye) - np.kron(eye, H.T))   # noqa: E731
    L0 = comm(H_list[0])
    for c in c_ops:
        cdc = c.conj().T @ c
        L0 = L0 + np.kron(c, c.conj()) - 0.5 * (np.kron(cdc, eye) + np.kron(eye, cdc.T))
    return [L0] + [comm(h) for h, _ in H_list[1:]]


def midpoint_coefficients(time, pulses, substeps):
    """Cubic-spline samples of the pulse arrays at the midpoints of `substeps` equal parts of every interval."""
    time = np.asarray(time, dtype=np.float64)
    edges = np.concatenate([np.linspace(time[k], time[k + 1], substeps + 1)[:-1] for k in range(len(time) - 1)]
                           + [time[-1:]])
    mids = 0.5 * (edges[:-1] + edges[1:])
    cols = [np.ones_like(mids)]
    for p in pulses:
        cols.append(CubicSpline(time, np.asarray(p, dtype=np.float64))(mids) if not callable(p)
                    else np.array([p(t, {}) for t in mids]))
    return mids, np.stack(cols, axis=1)


def default_substeps(L_pieces, time, pulses):
    """Smallest sub-stepping with |L dt|_1 <= 0.5 (entrywise bound at the largest sampled pulse amplitudes)."""
    peak = [1.0] + [float(np.max(np.abs(np.asarray(p)))) if not callable(p) else 1.0 for p in pulses]
    bound = sum(a * np.abs(L).sum(axis=0) for a, L in zip(peak, L_pieces)).max()
    dt = float(np.max(np.diff(time)))
    return max(1, int(np.ceil(bound * dt / 0.5)))


class Result:
    """The fields of qutip's Result the notebooks read (NB05 cells 21-22)."""

    def __init__(self, times, expect, final_state):
        self.times, self.expect, self.states, self.final_state = times, expect, [final_state], final_state


def Mesolve_5lvl_t(time, argsME5lvl, output_states=False, options=None, substeps=None, device=None):
    """QF:88-142 / QS:57-103: evolve psi0 under the 5-level two-atom master equation over `time`.
    ``output_states=True`` tracks the populations of psi0 and psi_orth (QF:131) -- plus, as QS:93-96, of |i> on
    atom 1 and of |rr> -- in ``result.expect`` (one real array per observable, one value per entry of `time`).
    ``result.final_state`` is the 25 x 25 density matrix at the last time.  ``options`` is accepted for signature
    compatibility and ignored (QuTiP's ODE controls do not apply)."""
    from qocvl import lindblad
    psi0, psi_orth = _ket(argsME5lvl[0]), _ket(argsME5lvl[1])
    Gammas = argsME5lvl[12]
    H = QTHamiltonian(argsME5lvl)
    pieces = liouvillian_pieces(H, collapse_operators(Gammas))
    pulses = [f for _, f in H[1:]]
    time = np.asarray(time, dtype=np.float64)
    if time.ndim != 1 or len(time) < 2 or np.any(np.diff(time) <= 0) or not np.allclose(np.diff(time), time[1] - time[0]):
        raise ValueError("time must be an increasing, uniformly spaced 1-D array")
    for p in pulses:
        if not callable(p) and np.shape(p) != time.shape:
            raise ValueError("pulse sample arrays must have the shape of `time`")
    sub = default_substeps(pieces, time, pulses) if substeps is None else int(substeps)
    _, coef = midpoint_coefficients(time, pulses, sub)
    rho0 = np.outer(psi0, psi0.conj()).reshape(-1)
    obs = []
    if output_states:
        sig_ii, sig_rr, eye = _op(2, 2), _op(3, 3), np.eye(DIM)
        for o in (np.outer(psi0, psi0.conj()), np.outer(psi_orth, psi_orth.conj()), np.kron(sig_ii, eye),
                  np.kron(sig_rr, sig_rr)):
            obs.append(o.T.reshape(-1))                       # Tr(O rho) = sum_ab O[b,a] rho[a,b]
    expect, rho = lindblad.propagate(pieces, coef, rho0, (time[1] - time[0]) / sub,
                                     obs=np.array(obs) if obs else None, device=device)
    series = [expect[0, ::sub, o].real.copy() for o in range(len(obs))]
    return Result(time, series, rho[0].reshape(DIM * DIM, DIM * DIM))
