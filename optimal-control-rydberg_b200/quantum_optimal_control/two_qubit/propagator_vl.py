"""Two blockaded atoms (4 levels each), global CZ gate: B200 mirror of the reference class
``two_qubit/propagator_vl.py::PropagatorVL`` (TQ:48-574) and of ``single_optimization_step``
(TQ:580-600).  Same constructor, attributes and method names; the reference's unused
``arc`` / ``qutip`` imports (TQ:7-8) are not needed."""
from __future__ import annotations

import numpy as np

import qocvl
from .._engine import EngineBase, block_diag2, block_upper, gradient_step, halving_flags

L0, L1, LI, LR = 0, 1, 2, 3        # single-atom level order (TQ:105-108)


def _op(row, col, dim=4):
    m = np.zeros((dim, dim), dtype=np.complex128)
    m[row, col] = 1.0
    return m


def _both_atoms(single):
    """TQ:201-207: I (x) h + h (x) I."""
    eye = np.eye(4, dtype=np.complex128)
    return np.kron(eye, single) + np.kron(single, eye)


def filter_impulse_response(no_of_steps, pad, f_std, delta_t):
    """Real taps h[M] of the frequency filter TQ:96-102 + TQ:375-397 (everything in it is linear:
    zero-pad, fft, fftshift, Gaussian window, ifftshift, ifft, real part), so that
    wave[k] = sum_t h[(k - t - pad) mod M] * reg[t].  Returns (window, taps)."""
    M = no_of_steps + 2 * pad
    freqs = np.fft.fftfreq(no_of_steps, d=delta_t)          # width uses no_of_steps, not M (quirk q5)
    std = f_std / (freqs[1] - freqs[0])
    x = np.arange(M, dtype=np.float64)
    window = np.exp(-0.5 * ((x - 0.5 * (M - 1)) / std) ** 2).astype(np.complex128)
    delta = np.zeros(M, dtype=np.complex128)
    delta[0] = 1.0
    taps = np.fft.ifft(np.fft.ifftshift(window * np.fft.fftshift(np.fft.fft(delta)))).real
    return window, taps


class PropagatorVL(EngineBase):
    _model = qocvl.MODEL_TWO_QUBIT
    W_INFID, W_ADIAB = 0.2, 0.8                    # TQ:574

    def __init__(self, input_dim: int, no_of_steps: int, pad: int, f_std: float, delta_t: float,
                 del_total: float, V_int: float, Delta_i: float, Rabi_i: float, Rabi_r: float,
                 Gammas_all: list, *, device=None):
        self.dim = 4
        self.input_dim = input_dim
        self.no_of_steps = no_of_steps
        self.padding = pad
        self.delta_t = delta_t
        self.del_total = del_total
        self.duration = (no_of_steps + 2 * pad) * delta_t       # TQ:88
        self.Vint = V_int
        self.Rabi_i_val, self.Rabi_r_val, self.Delta_i_val = Rabi_i, Rabi_r, Delta_i
        self.Gamma_10, self.Gamma_i1, self.Gamma_ri, self.Gamma_rd = Gammas_all
        M = no_of_steps + 2 * pad
        self.Gaussian_filter, taps = filter_impulse_response(no_of_steps, pad, f_std, delta_t)
        one = np.eye(4, dtype=np.complex128)
        g0, g1 = one[L0], one[L1]
        self.psi_10, self.psi_11, self.psi_01 = np.kron(g1, g0), np.kron(g1, g1), np.kron(g0, g1)
        self.psi_pp = 0.5 * (np.kron(g0, g0) + np.kron(g0, g1) + np.kron(g1, g0) + np.kron(g1, g1))
        signs = np.zeros(16)
        signs[[0, 1, 4]] = 1.0                                   # |00>, |01>, |10>
        signs[5] = -1.0                                          # |11>        (TQ:121-127)
        self.U_target = np.diag(signs).astype(np.complex128)
        self.contraction_array = halving_flags(M)                # TQ:139
        pieces = self._Hamiltonian(g0, g1, one[LI], one[LR], del_total, V_int, Delta_i, Rabi_i, Rabi_r,
                                   *Gammas_all)
        self.VL_Generators_ansatz = self._VL_Generators(pieces, g0, g1)
        # engine generators: unit-strength total detuning so a noise ensemble can move it per sample;
        # TQ:470 multiplies the two drift pieces by -1j.
        unit = self._Hamiltonian(g0, g1, one[LI], one[LR], 1.0, V_int, Delta_i, Rabi_i, Rabi_r, *Gammas_all)
        p00 = _op(L0, L0)
        tops = [np.kron(_op(L1, L1), p00), np.kron(_op(LR, LR), p00),
                np.kron(_op(L1, LR), p00), np.kron(_op(LR, L1), p00)]          # TQ:303-307
        gens_u = np.stack(unit + [np.zeros((16, 16), complex)] * 4)
        gens_w = np.stack([np.zeros((16, 16), complex)] * 7 + tops)
        coef_const = np.zeros(11, complex)
        coef_const[0], coef_const[1] = -1j, -1j * del_total
        cols = [int(c) for c in np.flatnonzero(signs)]           # columns of U_target the trace reads
        eye16 = np.eye(16, dtype=np.complex128)
        self._make_plan(n=16, n_slices=M, n_basis_pts=no_of_steps, pad=pad, gens_u=gens_u, gens_w=gens_w,
                        coef_const=coef_const, starts=[eye16[c] for c in cols],
                        targets=[self.U_target[:, c] for c in cols], adiab_index=cols.index(4),
                        fid_norm=float(np.sum(np.abs(signs)) ** 2),          # (Tr Ut^dag Ut)^2 = 16, TQ:555
                        scale=(Rabi_i, Rabi_r, 0, 0), taps=taps, device=device,
                        # both atoms see the same global drive (TQ:201-207): exchanging them is a symmetry of
                        # every Hamiltonian piece, so U|01> is the mirror image of U|10> (engine-verified hint)
                        symmetry=[4 * (s % 4) + s // 4 for s in range(16)])

    gen_contraction_array = staticmethod(halving_flags)

    @staticmethod
    def _Hamiltonian(g_0, g_1, i_st, r_st, del_total, V_int, Delta_i, Rabi_i, Rabi_r,
                     Gamma_10, Gamma_i1, Gamma_ri, Gamma_rd):
        """TQ:159-217: seven 16x16 pieces [H0, del_total, Delta, Rabi_i re/im, Rabi_r re/im]."""
        decay = -0.5j * (Gamma_i1 * _op(LI, LI) + Gamma_10 * _op(L1, L1) + (Gamma_ri + Gamma_rd) * _op(LR, LR))
        one_i, i_r = Rabi_i * _op(L1, LI), Rabi_r * _op(LI, LR)
        return [_both_atoms(decay) + V_int * np.kron(_op(LR, LR), _op(LR, LR)),
                _both_atoms(-del_total * _op(LR, LR)),
                _both_atoms(-Delta_i * _op(LI, LI)),
                _both_atoms(-0.5 * (one_i + one_i.T)), _both_atoms(-0.5j * (one_i - one_i.T)),
                _both_atoms(-0.5 * (i_r + i_r.T)), _both_atoms(-0.5j * (i_r - i_r.T))]

    @staticmethod
    def _VL_Generators(Hams, g_0, g_1):
        """TQ:219-314: dense [11, 32, 32] stack (attribute kept for callers)."""
        p00 = _op(L0, L0)
        tops = [np.kron(_op(L1, L1), p00), np.kron(_op(LR, LR), p00),
                np.kron(_op(L1, LR), p00), np.kron(_op(LR, L1), p00)]
        return np.stack([block_diag2(np.asarray(h, complex)) for h in Hams] + [block_upper(t) for t in tops])

    def _noise_tables(self, del_total, rabi_scale):
        S = len(del_total)
        mul = np.ones((S, 11))
        mul[:, 3:7] = rabi_scale[:, None]
        add = np.zeros((S, 11), complex)
        add[:, 1] = -1j * (del_total - self.del_total)
        return mul, add


def single_optimization_step(propagator: PropagatorVL, ctrl_a, ctrl_b, ctrl_c, lr: float = 0.02):
    """TQ:580-600: one plain gradient-descent update -> (cost, new_a, new_b, new_c)."""
    return gradient_step(propagator, ctrl_a, ctrl_b, ctrl_c, lr)
