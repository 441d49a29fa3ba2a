"""NumPy/SciPy restatement of the two-atom global-CZ engine.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
/root/reference/code/quantum_optimal_control/two_qubit/propagator_vl.py ("TQ").
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .state_transfer import (C128, gaussian_modes, halving_flags, ketbra, lift_diag,
                             lift_upper, ordered_product, unit_ket)


def window_gaussian(length: int, std: float) -> np.ndarray:
    """TQ:35-42 -- exp(-0.5((x-(L-1)/2)/std)^2)."""
    x = np.arange(length, dtype=np.float64)
    return np.exp(-0.5 * ((x - 0.5 * (length - 1)) / std) ** 2)


class PropagatorVL:
    """TQ:48-574."""

    W_INFID = 0.2  # TQ:574
    W_ADIAB = 0.8
    EPS = 1e-12    # TQ:433

    def __init__(self, input_dim, no_of_steps, pad, f_std, delta_t, del_total, V_int,
                 Delta_i, Rabi_i, Rabi_r, Gammas_all):
        # TQ:82-94
        self.dim = 4
        self.input_dim = input_dim
        self.no_of_steps = no_of_steps
        self.padding = pad
        self.delta_t = delta_t
        self.del_total = del_total
        self.duration = (no_of_steps + 2 * pad) * delta_t
        self.Vint = V_int
        self.Rabi_i_val, self.Rabi_r_val, self.Delta_i_val = Rabi_i, Rabi_r, Delta_i
        self.Gamma_10, self.Gamma_i1, self.Gamma_ri, self.Gamma_rd = Gammas_all
        # TQ:96-102 -- filter width in bins of fftfreq(no_of_steps), window over M points
        n_slices = no_of_steps + 2 * pad
        freqs = np.fft.fftfreq(no_of_steps, d=delta_t)
        self.Gaussian_filter = window_gaussian(n_slices, f_std / (freqs[1] - freqs[0])).astype(C128)
        # TQ:104-127
        g0, g1, i_, r = (unit_ket(4, k) for k in range(4))
        self.psi_10 = np.kron(g1, g0)
        self.psi_11 = np.kron(g1, g1)
        self.psi_01 = np.kron(g0, g1)
        self.psi_pp = 0.5 * (np.kron(g0, g0) + np.kron(g0, g1) + np.kron(g1, g0) + np.kron(g1, g1))
        self.U_target = (ketbra(np.kron(g0, g0)) + ketbra(np.kron(g0, g1))
                         + ketbra(np.kron(g1, g0)) - ketbra(np.kron(g1, g1)))
        pieces = self._Hamiltonian(g0, g1, i_, r, del_total, V_int, Delta_i, Rabi_i, Rabi_r,
                                   self.Gamma_10, self.Gamma_i1, self.Gamma_ri, self.Gamma_rd)
        self.VL_Generators_ansatz = self._VL_Generators(pieces, g0, g1)
        self.contraction_array = halving_flags(n_slices)

    gen_contraction_array = staticmethod(halving_flags)

    @staticmethod
    def _Hamiltonian(g0, g1, i_, r, del_total, V_int, Delta_i, Rabi_i, Rabi_r,
                     Gamma_10, Gamma_i1, Gamma_ri, Gamma_rd):
        """TQ:159-217 -- seven 16x16 pieces, each single-atom piece summed over both atoms."""
        s11, sii, srr = ketbra(g1), ketbra(i_), ketbra(r)
        s1i, sir = ketbra(g1, i_), ketbra(i_, r)
        decay = -0.5j * Gamma_i1 * sii - 0.5j * Gamma_10 * s11 - 0.5j * (Gamma_ri + Gamma_rd) * srr
        h_del = -del_total * srr
        h_Delta = -Delta_i * sii

        def herm_part(op):   # TQ:185-186
            return -0.5 * (op + op.conj().T)

        def anti_part(op):   # TQ:187-188
            return -0.5j * (op - op.conj().T)

        eye = np.eye(4, dtype=C128)

        def both(op):        # TQ:202-207  I (x) op + op (x) I  (order irrelevant for a sum)
            return np.kron(eye, op) + np.kron(op, eye)

        return [both(decay) + V_int * np.kron(srr, srr), both(h_del), both(h_Delta),
                both(herm_part(Rabi_i * s1i)), both(anti_part(Rabi_i * s1i)),
                both(herm_part(Rabi_r * sir)), both(anti_part(Rabi_r * sir))]

    @staticmethod
    def _VL_Generators(pieces, g0, g1):
        """TQ:219-314 -- the live projector definitions are TQ:303-307."""
        r = unit_ket(4, 3)
        p00 = ketbra(g0)
        uppers = [np.kron(ketbra(g1), p00), np.kron(ketbra(r), p00),
                  np.kron(ketbra(g1, r), p00), np.kron(ketbra(r, g1), p00)]
        return np.stack([lift_diag(p.astype(C128)) for p in pieces]
                        + [lift_upper(u.astype(C128)) for u in uppers], axis=0)

    # ---- pulse map -------------------------------------------------------
    def _transform_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:335-353."""
        cols = []
        for j in range(ctrl_a.shape[1]):
            modes = gaussian_modes(self.no_of_steps, np.tanh(ctrl_b[:, j]), np.tanh(ctrl_c[:, j]))
            cols.append(modes @ ctrl_a[:, j])
        return np.stack(cols, axis=1)

    def _regularize_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:355-373 -- detuning channel forced to -1 (quirk q2)."""
        raw = self._transform_amplitudes(ctrl_a, ctrl_b, ctrl_c)
        return np.concatenate([0.0 * raw[:, 0:1] - 1.0,
                               1.0 - np.exp(-np.square(raw[:, 1:2])), np.tanh(raw[:, 2:3]),
                               1.0 - np.exp(-np.square(raw[:, 3:4])), np.tanh(raw[:, 4:5])], axis=1)

    def _filter_amplitudes(self, amplitudes):
        """TQ:375-397 -- zero-pad, fft, fftshift, window, ifftshift, ifft, real part."""
        pad = self.padding
        padded = np.pad(amplitudes.astype(C128), ((pad, pad), (0, 0)), mode="constant")
        spectrum = np.fft.fftshift(np.fft.fft(padded, axis=0), axes=0)
        back = np.fft.ifft(np.fft.ifftshift(self.Gaussian_filter[:, None] * spectrum, axes=0), axis=0)
        return np.real(back)

    def return_physical_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:399-404."""
        return self._filter_amplitudes(self._regularize_amplitudes(
            np.asarray(ctrl_a, float), np.asarray(ctrl_b, float), np.asarray(ctrl_c, float)))

    def aux_from_wave(self, wave):
        """TQ:416-455."""
        om_i = (self.Rabi_i_val * wave[:, 1]) * np.exp(1j * (np.pi * wave[:, 2]))
        om_r = (self.Rabi_r_val * wave[:, 3]) * np.exp(1j * (np.pi * wave[:, 4]))
        p1 = om_r * np.conj(om_r)
        p2 = om_i * np.conj(om_i)
        p3 = -(om_i * om_r)
        p4 = -(np.conj(om_i) * np.conj(om_r))
        denom = p1 + p2
        # jnp "<" on complex is lexicographic; the imaginary part is exactly 0 here
        denom = np.where(denom.real < self.EPS, self.EPS, denom)
        weights = np.stack([p1 / denom, p2 / denom, p3 / denom, p4 / denom], axis=1)
        ham = np.stack([wave[:, 0], om_i.real / self.Rabi_i_val, om_i.imag / self.Rabi_i_val,
                        om_r.real / self.Rabi_r_val, om_r.imag / self.Rabi_r_val], axis=1).astype(C128)
        return np.concatenate([-1j * ham, weights], axis=1)

    def return_auxiliary_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:406-455."""
        return self.aux_from_wave(self.return_physical_amplitudes(ctrl_a, ctrl_b, ctrl_c))

    # ---- propagators -----------------------------------------------------
    def exponent(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:466-476 -- drift IS multiplied by -1j here."""
        gens = self.VL_Generators_ansatz
        aux = self.return_auxiliary_amplitudes(ctrl_a, ctrl_b, ctrl_c)
        drift = -1j * (gens[0] + gens[1])
        return self.delta_t * (drift[None] + np.tensordot(aux, gens[2:], axes=[[1], [0]]))

    def exponentials(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:460-484."""
        return sla.expm(self.exponent(ctrl_a, ctrl_b, ctrl_c))

    def propagate(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:489-535."""
        return ordered_product(self.exponentials(ctrl_a, ctrl_b, ctrl_c), self.contraction_array)

    def metrics_of(self, vl):
        """TQ:549-567 applied to a given Van Loan propagator."""
        n = self.dim * self.dim
        u = vl[:n, :n]
        tr = np.trace(self.U_target.conj().T @ u)
        tr_norm = np.trace(self.U_target.conj().T @ self.U_target)
        infidelity = 1.0 - np.real((tr * np.conj(tr)) / tr_norm ** 2)
        final = u @ self.psi_10
        overlap = np.conj(final) @ (vl[:n, n:2 * n] @ self.psi_10)
        adiabaticity = np.real(1.0 - (1.0 / self.duration) * overlap)
        return float(infidelity), float(adiabaticity)

    def metrics(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:540-567."""
        return self.metrics_of(self.propagate(ctrl_a, ctrl_b, ctrl_c))

    def target(self, ctrl_a, ctrl_b, ctrl_c):
        """TQ:569-574."""
        infid, adiab = self.metrics(ctrl_a, ctrl_b, ctrl_c)
        return self.W_INFID * infid + self.W_ADIAB * adiab


# fixed physical constants for the synthetic configs (ARC not needed)
GAMMAS_CZ = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]


def notebook02_setup(seed=0, n_gauss=10, nt=500, del_total=0.0, rabi_scale=1.0, total_time=None):
    """cfg 2 of SURVEY.md 8(d): notebook 02 cell 4 parameters; a,b,c from default_rng(seed)
    (the notebook's jax.random.PRNGKey(6) draw is not reproducible without jax).
    With nt=1888, n_gauss=16 this is the per-sample problem of cfg 3."""
    pad = int(0.03 * nt)
    if total_time is None:   # reduced-size test grids keep the notebook's time step (648 ns / 530)
        total_time = 648e-9 if nt >= 500 else 648e-9 * (nt + 2 * pad) / 530
    dt = total_time / (nt + 2 * pad)
    prop = PropagatorVL(n_gauss, nt, pad, 50e6, dt, del_total, 2 * np.pi * 10e6,
                        2 * np.pi * (-35.7e6), 2 * np.pi * 100e6 * rabi_scale,
                        2 * np.pi * 100e6 * rabi_scale, GAMMAS_CZ)
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (n_gauss, 5))
    b = rng.uniform(-1, 1, (n_gauss, 5))
    c = rng.uniform(0, 0.5, (n_gauss, 5))
    return prop, a, b, c
