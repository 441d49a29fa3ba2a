"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's master-equation validation
(two_qubit/qutip_five_level.py:11-142, two_qubit/qutip_simulation.py:7-103): two 5-level atoms (g0, g1, i, r, dark),
the CZ Hamiltonian pieces with time-dependent pulse coefficients, 8 collapse operators, Lindblad equation.

QuTiP itself (the third-party solver the reference calls, version unpinned, not installed here) is replaced by two
independent integrators of the same equation:
  * ``evolve_midpoint``: exponential midpoint rule with scipy.sparse.linalg.expm_multiply -- the SAME discretisation
    the GPU kernel uses, but with SciPy's exponential action instead of the kernel's Taylor series;
  * ``evolve_ode``: scipy.integrate.solve_ivp (DOP853, tight tolerances) on the cubic-spline-interpolated pulses, the
    way qutip.mesolve treats array coefficients -- an independent check of the discretisation itself.
Parity is unpinned by the reference (no stored mesolve outputs); the two integrators pin each other.
"""
import numpy as np
import scipy.sparse as sp
from scipy.integrate import solve_ivp
from scipy.interpolate import CubicSpline
from scipy.sparse.linalg import expm_multiply

DIM = 5
G0, G1, LI, LR, LD = range(5)


def ket(level):
    v = np.zeros(DIM, complex)
    v[level] = 1.0
    return v


def sigma(a, b):
    """|a><b| on one atom."""
    return np.outer(ket(a), ket(b).conj())


def two_atom(single):
    eye = np.eye(DIM)
    return np.kron(single, eye) + np.kron(eye, single)


def hamiltonian_pieces(V_int, Rabi_i, Rabi_r, Delta_i, del_total):
    """QF:35-75: static part and the five driven parts (Delta_i, Rabi_i re/im, Rabi_r re/im)."""
    s1i, sir = sigma(G1, LI), sigma(LI, LR)
    H0 = two_atom(-del_total * sigma(LR, LR)) + V_int * np.kron(sigma(LR, LR), sigma(LR, LR))
    return [H0,
            two_atom(-Delta_i * sigma(LI, LI)),
            two_atom(-0.5 * Rabi_i * (s1i + s1i.conj().T)),
            two_atom(-0.5j * Rabi_i * (s1i - s1i.conj().T)),
            two_atom(-0.5 * Rabi_r * (sir + sir.conj().T)),
            two_atom(-0.5j * Rabi_r * (sir - sir.conj().T))]


def collapse_ops(Gammas):
    """QF:104-118: [Gamma_10, Gamma_i1, Gamma_ri, Gamma_rd] on either atom."""
    eye = np.eye(DIM)
    out = []
    for gamma, s in zip(Gammas, (sigma(G0, G1), sigma(G1, LI), sigma(LI, LR), sigma(LD, LR))):
        out += [np.sqrt(gamma) * np.kron(s, eye), np.sqrt(gamma) * np.kron(eye, s)]
    return out


def rhs(rho, H, c_ops):
    """Lindblad right-hand side on a 25 x 25 density matrix (no vectorisation: independent of the kron algebra)."""
    out = -1j * (H @ rho - rho @ H)
    for c in c_ops:
        cd = c.conj().T
        out += c @ rho @ cd - 0.5 * (cd @ c @ rho + rho @ cd @ c)
    return out


def superoperators(H_pieces, c_ops):
    """Row-major vec: vec(A rho B) = (A kron B^T) vec(rho)."""
    n = H_pieces[0].shape[0]
    eye = np.eye(n)
    comm = [-1j * (np.kron(H, eye) - np.kron(eye, H.T)) for H in H_pieces]
    for c in c_ops:
        cdc = c.conj().T @ c
        comm[0] = comm[0] + np.kron(c, c.conj()) - 0.5 * (np.kron(cdc, eye) + np.kron(eye, cdc.T))
    return comm


def midpoint_times(time, substeps):
    edges = np.concatenate([np.linspace(time[k], time[k + 1], substeps + 1)[:-1] for k in range(len(time) - 1)]
                           + [time[-1:]])
    return 0.5 * (edges[:-1] + edges[1:])


def evolve_midpoint(time, pulses, H_pieces, c_ops, rho0, substeps):
    """rho at every entry of `time` ([len(time), 25, 25]); pulses: 5 sample arrays on `time`."""
    L = [sp.csr_matrix(m) for m in superoperators(H_pieces, c_ops)]
    mids = midpoint_times(time, substeps)
    coef = np.stack([np.ones_like(mids)] + [CubicSpline(time, p)(mids) for p in pulses], axis=1)
    dt = (time[1] - time[0]) / substeps
    v = rho0.reshape(-1).astype(complex)
    out = [v.copy()]
    for k in range(len(mids)):
        Lk = sum(c * m for c, m in zip(coef[k], L))
        v = expm_multiply(Lk * dt, v)
        if (k + 1) % substeps == 0:
            out.append(v.copy())
    return np.array(out).reshape(len(time), DIM * DIM, DIM * DIM)


def evolve_ode(time, pulses, H_pieces, c_ops, rho0, rtol=1e-10, atol=1e-12):
    splines = [CubicSpline(time, p) for p in pulses]
    n = DIM * DIM

    def f(t, y):
        H = H_pieces[0] + sum(s(t) * h for s, h in zip(splines, H_pieces[1:]))
        return rhs(y.reshape(n, n), H, c_ops).reshape(-1)

    sol = solve_ivp(f, (time[0], time[-1]), rho0.reshape(-1).astype(complex), method="DOP853", t_eval=time,
                    rtol=rtol, atol=atol)
    return sol.y.T.reshape(len(time), n, n)
