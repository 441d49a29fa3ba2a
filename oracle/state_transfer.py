"""NumPy/SciPy restatement of the single-atom (5-level) state-transfer engine.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, function by function,
/root/reference/code/quantum_optimal_control/state_transfer/propagator_vl.py ("ST").
Every method cites the ST lines it restates.  All arithmetic is float64/complex128.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

C128 = np.complex128


def unit_ket(dim: int, idx: int) -> np.ndarray:
    """ST:14-17 -- |idx> in a dim-level space."""
    ket = np.zeros(dim, dtype=C128)
    ket[idx] = 1.0
    return ket


def ketbra(left: np.ndarray, right: np.ndarray | None = None) -> np.ndarray:
    """ST:19-21 (and the explicit jnp.outer(x, conj(y)) calls) -- |left><right|."""
    right = left if right is None else right
    return np.outer(left, np.conj(right))


def halving_flags(count: int) -> list[bool]:
    """ST:77-87 -- odd/even flag per level of the pairwise product tree."""
    flags = []
    while count > 1:
        flags.append(bool(count % 2))
        count //= 2
    return flags


def lift_diag(block: np.ndarray) -> np.ndarray:
    """[[block,0],[0,block]] -- ST:146-159."""
    zero = np.zeros_like(block)
    return np.block([[block, zero], [zero, block]])


def lift_upper(block: np.ndarray) -> np.ndarray:
    """[[0,block],[0,0]] -- ST:161-168."""
    zero = np.zeros_like(block)
    return np.block([[zero, block], [zero, zero]])


def ordered_product(step_ops: np.ndarray, flags: list[bool]) -> np.ndarray:
    """ST:258-283 / TQ:489-535 -- later slice on the LEFT, odd leftover folded into the tail."""
    ops = step_ops
    for odd in flags:
        if odd:
            leftover = ops[-1]
            ops = ops[:-1]
            merged = np.matmul(ops[1::2], ops[0:-1:2])
            tail = leftover @ merged[-1]
            ops = np.concatenate([merged[:-1], tail[None]], axis=0)
        else:
            ops = np.matmul(ops[1::2], ops[0::2])
    return np.squeeze(ops)


def gaussian_modes(n_pts: int, centres: np.ndarray, widths: np.ndarray) -> np.ndarray:
    """ST:173-185 -- exp(-(t-b)^2/c^2) on t=linspace(-1,1,n_pts); returns [n_pts, N]."""
    t = np.linspace(-1.0, 1.0, n_pts)
    return np.exp(-((t[:, None] - centres[None, :]) ** 2) / (widths[None, :] ** 2))


class PropagatorVL:
    """ST:32-310."""

    W_INFID = 0.6  # ST:310
    W_ADIAB = 0.4

    def __init__(self, input_dim, no_of_steps, delta_t, del_total, Delta_1,
                 Rabi_1, Rabi_2, Gamma_r, Gamma_i):
        # ST:46-56
        self.duration = no_of_steps * delta_t
        self.delta_t = delta_t
        self.Rabi_1, self.Rabi_2 = Rabi_1, Rabi_2
        self.Delta_1, self.del_total = Delta_1, del_total
        self.Gamma_r, self.Gamma_i = Gamma_r, Gamma_i
        self.dim = 5
        self.input_dim = input_dim
        self.no_of_steps = no_of_steps
        kets = self.nLevelAtomBasis(self.dim)
        self.psi_0 = kets[2]  # |g>   ST:60
        self.psi_t = kets[4]  # |r>   ST:61
        self.VL_Generators_ansatz = self.VL_Generators(
            self.Hamiltonian(kets + [del_total, Delta_1, Rabi_1, Rabi_2, Gamma_r, Gamma_i]),
            kets, self.dim)
        self.num_ctrls = 5
        self.contraction_array = halving_flags(no_of_steps)

    gen_contraction_array = staticmethod(halving_flags)

    @staticmethod
    def nLevelAtomBasis(n):
        """ST:89-95."""
        return [unit_ket(n, k) for k in range(n)]

    @staticmethod
    def Hamiltonian(args):
        """ST:97-130 -- the seven 5x5 pieces (order: decay, del_total, Delta_1, 1re, 1im, 2re, 2im)."""
        gp, rp, g, i_, r, del_total, Delta_1, Rabi_1, Rabi_2, Gamma_r, Gamma_i = args
        s_ii, s_rr = ketbra(i_), ketbra(r)
        s_gi, s_ir = ketbra(g, i_), ketbra(i_, r)
        decay = -0.5j * Gamma_i * s_ii - 0.5j * Gamma_r * s_rr
        h_del = -del_total * s_rr
        h_Delta = -Delta_1 * s_ii
        h1_re = -0.5 * Rabi_1 * (s_gi + s_gi.conj().T)
        h1_im = -0.5j * Rabi_1 * (s_gi - s_gi.conj().T)
        h2_re = -0.5 * Rabi_2 * (s_ir + s_ir.conj().T)
        h2_im = -0.5j * Rabi_2 * (s_ir - s_ir.conj().T)
        return [decay, h_del, h_Delta, h1_re, h1_im, h2_re, h2_im]

    @staticmethod
    def VL_Generators(pieces, kets, dim):
        """ST:132-170 -- 7 diagonal lifts + 4 projector pieces in the top-right block."""
        gp, rp, g, i_, r = kets
        lifted = [lift_diag(p.astype(C128)) for p in pieces]
        lifted += [lift_upper(ketbra(g)), lift_upper(ketbra(r)),
                   lift_upper(ketbra(g, r)), lift_upper(ketbra(r, g))]
        return np.stack(lifted, axis=0)

    # ---- pulse map -------------------------------------------------------
    def transform_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        """ST:187-202."""
        centres, widths = np.tanh(ctrl_b), np.tanh(ctrl_c)
        cols = [gaussian_modes(self.no_of_steps, centres[:, j], widths[:, j]) @ ctrl_a[:, j]
                for j in range(self.num_ctrls)]
        return np.stack(cols, axis=1)

    @staticmethod
    def regularize_amplitudes(pair):
        """ST:204-213 -- radial squash tanh(r)/r of an (re, im) pair."""
        radius = np.sqrt(pair[:, 0] ** 2 + pair[:, 1] ** 2)
        with np.errstate(invalid="ignore", divide="ignore"):
            scale = np.where(radius == 0, 1.0, np.tanh(radius) / radius)
        return scale[:, None] * pair

    def return_physical_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        """ST:215-227."""
        raw = self.transform_amplitudes(np.asarray(ctrl_a, float), np.asarray(ctrl_b, float),
                                        np.asarray(ctrl_c, float))
        return np.concatenate([np.tanh(raw[:, 0:1]),
                               self.regularize_amplitudes(raw[:, 1:3]),
                               self.regularize_amplitudes(raw[:, 3:5])], axis=1)

    def aux_from_wave(self, wave):
        """ST:236-245 (everything after the pulse map)."""
        om1 = self.Rabi_1 * (wave[:, 1:2] + 1j * wave[:, 2:3])
        om2 = self.Rabi_2 * (wave[:, 3:4] + 1j * wave[:, 4:5])
        p1 = om2 * np.conj(om2)
        p2 = om1 * np.conj(om1)
        p3 = -om1 * om2
        p4 = -np.conj(om1) * np.conj(om2)
        with np.errstate(invalid="ignore", divide="ignore"):
            weights = np.concatenate([p1, p2, p3, p4], axis=1) / (p1 + p2)
        return np.concatenate([-1j * wave.astype(C128), weights.astype(C128)], axis=1)

    def return_auxiliary_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        """ST:229-245."""
        return self.aux_from_wave(self.return_physical_amplitudes(ctrl_a, ctrl_b, ctrl_c))

    # ---- propagators -----------------------------------------------------
    def exponent(self, ctrl_a, ctrl_b, ctrl_c):
        """ST:252-255 -- NOTE the drift (gens 0,1) is NOT multiplied by -1j (quirk q1)."""
        gens = self.VL_Generators_ansatz
        aux = self.return_auxiliary_amplitudes(ctrl_a, ctrl_b, ctrl_c)
        return self.delta_t * (gens[0] + gens[1] + np.tensordot(aux, gens[2:], axes=[[1], [0]]))

    def exponentials(self, ctrl_a, ctrl_b, ctrl_c):
        """ST:247-256."""
        return sla.expm(self.exponent(ctrl_a, ctrl_b, ctrl_c))

    def propagate(self, ctrl_a, ctrl_b, ctrl_c):
        """ST:258-283."""
        return ordered_product(self.exponentials(ctrl_a, ctrl_b, ctrl_c), self.contraction_array)

    def metrics_of(self, vl):
        """ST:291-302 applied to a given Van Loan propagator."""
        n = self.dim
        final = vl[:n, :n] @ self.psi_0
        overlap = np.conj(self.psi_t) @ final
        infidelity = 1 - np.real(np.conj(overlap) * overlap)
        inner = vl[:n, n:2 * n] @ self.psi_0
        adiabaticity = np.real(1 - (1 / self.duration) * (np.conj(final) @ inner))
        return float(infidelity), float(adiabaticity)

    def metrics(self, ctrl_a, ctrl_b, ctrl_c):
        """ST:285-302."""
        return self.metrics_of(self.propagate(ctrl_a, ctrl_b, ctrl_c))

    def target(self, ctrl_a, ctrl_b, ctrl_c):
        """ST:304-310."""
        infid, adiab = self.metrics(ctrl_a, ctrl_b, ctrl_c)
        return self.W_INFID * infid + self.W_ADIAB * adiab


def notebook01_setup(Gamma_i=1.0 / 118e-9, Gamma_r=1.0 / 150e-6):
    """Parameters and the np.random.seed(1) draw of notebook 01 cells 1, 3, 5.

    The two decay rates come from the ARC package in the notebook (absent here);
    SURVEY.md section 0.3: the cost moves by 2e-8 over 100-130 ns of 6P3/2 lifetime.
    Returns (propagator, ctrl_a, ctrl_b, ctrl_c).
    """
    rabi = 2 * np.pi * 127e6
    nt, n_gauss = 500, 10
    dt = 200e-9 / nt
    prop = PropagatorVL(n_gauss, nt, dt, 0, 0, rabi, rabi, Gamma_r, Gamma_i)
    rng = np.random.RandomState(1)  # == np.random.seed(1) global stream
    a = rng.uniform(-1, 1, (n_gauss, 5))
    b = np.concatenate([rng.uniform(-1, 1, (n_gauss, 1)),
                        rng.uniform(0, 1, (n_gauss, 2)),
                        rng.uniform(-1, 0, (n_gauss, 2))], axis=1)
    c = rng.uniform(0, 0.1, (n_gauss, 5))
    return prop, a, b, c


NOTEBOOK01_INITIAL_COST = 0.8022425853500806  # 01_state_transfer.ipynb cell 5 output
