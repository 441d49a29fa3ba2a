"""Single-atom (5-level) ground -> Rydberg state transfer: B200 mirror of the reference class
``state_transfer/propagator_vl.py::PropagatorVL`` (ST:32-310).  Same constructor, attributes and
method names; arrays in and out are float64 [input_dim, 5]; results are torch CUDA tensors."""
from __future__ import annotations

import numpy as np

import qocvl
from .._engine import EngineBase, block_diag2, block_upper, gradient_step, halving_flags

LEVELS = ("g_prime", "r_prime", "g", "i", "r")     # ST:59 basis order
G_, I_, R_ = 2, 3, 4


def _op(row, col, dim=5):
    m = np.zeros((dim, dim), dtype=np.complex128)
    m[row, col] = 1.0
    return m


class PropagatorVL(EngineBase):
    _model = qocvl.MODEL_STATE_TRANSFER
    W_INFID, W_ADIAB = 0.6, 0.4                    # ST:310

    def __init__(self, input_dim: int, no_of_steps: int, delta_t: float, del_total: float, Delta_1: float,
                 Rabi_1: float, Rabi_2: float, Gamma_r: float, Gamma_i: float, *, device=None):
        self.duration = no_of_steps * delta_t      # ST:46
        self.delta_t = delta_t
        self.Rabi_1, self.Rabi_2 = Rabi_1, Rabi_2
        self.Delta_1, self.del_total = Delta_1, del_total
        self.Gamma_r, self.Gamma_i = Gamma_r, Gamma_i
        self.dim = 5
        self.input_dim = input_dim
        self.no_of_steps = no_of_steps
        kets = self.nLevelAtomBasis(self.dim)
        self.psi_0, self.psi_t = kets[G_], kets[R_]             # ST:60-61
        self.contraction_array = halving_flags(no_of_steps)     # ST:75
        pieces = self.Hamiltonian([*kets, del_total, Delta_1, Rabi_1, Rabi_2, Gamma_r, Gamma_i])
        self.VL_Generators_ansatz = self.VL_Generators(pieces, kets, self.dim)
        # engine generators: same 11 pieces, but the total-detuning piece is kept with unit strength
        # so that a noise ensemble can move del_total per sample (coefficient = del_total, ST:254
        # adds the drift pieces with weight +1, NOT -1j).
        unit = self.Hamiltonian([*kets, 1.0, Delta_1, Rabi_1, Rabi_2, Gamma_r, Gamma_i])
        gens_u = np.stack(unit + [np.zeros((5, 5), complex)] * 4)
        gens_w = np.stack([np.zeros((5, 5), complex)] * 7
                          + [_op(G_, G_), _op(R_, R_), _op(G_, R_), _op(R_, G_)])   # ST:161-168
        coef_const = np.zeros(11, complex)
        coef_const[0], coef_const[1] = 1.0, del_total
        self._make_plan(n=5, n_slices=no_of_steps, n_basis_pts=no_of_steps, pad=0, gens_u=gens_u,
                        gens_w=gens_w, coef_const=coef_const, starts=[self.psi_0], targets=[self.psi_t],
                        adiab_index=0, fid_norm=1.0, scale=(Rabi_1, Rabi_2, 0, 0), taps=None, device=device)

    gen_contraction_array = staticmethod(halving_flags)

    @staticmethod
    def nLevelAtomBasis(n: int):
        """ST:89-95."""
        return [row.copy() for row in np.eye(n, dtype=np.complex128)]

    @staticmethod
    def Hamiltonian(args):
        """ST:97-130: [decay, total detuning, Delta_1, Rabi_1 re/im, Rabi_2 re/im] as 5x5 arrays."""
        *_, del_total, Delta_1, Rabi_1, Rabi_2, Gamma_r, Gamma_i = args
        gi, ir = _op(G_, I_), _op(I_, R_)
        return [-0.5j * (Gamma_i * _op(I_, I_) + Gamma_r * _op(R_, R_)),
                -del_total * _op(R_, R_),
                -Delta_1 * _op(I_, I_),
                -0.5 * Rabi_1 * (gi + gi.T), -0.5j * Rabi_1 * (gi - gi.T),
                -0.5 * Rabi_2 * (ir + ir.T), -0.5j * Rabi_2 * (ir - ir.T)]

    @staticmethod
    def VL_Generators(Hamiltonian, basis, dim: int):
        """ST:132-170: the dense [11, 10, 10] stack (kept for callers that read the attribute)."""
        tops = [_op(G_, G_, dim), _op(R_, R_, dim), _op(G_, R_, dim), _op(R_, G_, dim)]
        return np.stack([block_diag2(np.asarray(h, complex)) for h in Hamiltonian]
                        + [block_upper(t) for t in tops])

    def _noise_tables(self, del_total, rabi_scale):
        S = len(del_total)
        mul = np.ones((S, 11))
        mul[:, 3:7] = rabi_scale[:, None]           # Rabi_1, Rabi_2 pieces scale with the Rabi frequency
        add = np.zeros((S, 11), complex)
        add[:, 1] = del_total - self.del_total      # coefficient of the unit total-detuning piece
        return mul, add


def single_optimization_step(propagator, ctrl_a, ctrl_b, ctrl_c, lr: float = 0.1):
    """Notebook 01 cell 9 ``optimization_step``."""
    return gradient_step(propagator, ctrl_a, ctrl_b, ctrl_c, lr)
