"""Open chain of Rydberg atoms in the nearest-neighbour blockade (PXP) subspace: adiabatic preparation of the
Z2-ordered state.  This model is NOT in the reference; it is the extrapolation BASELINE.json configs 4-5 name
 written in the reference's pattern -- pulse-driven Hamiltonian pieces in the diagonal
blocks of the Van Loan generator, fixed projectors with time-dependent weights in the top-right block
(ST:132-170) -- and it reuses the same engine, cost definition (ST:285-310) and method surface.

  H(t) = Omega(t)/2 * sum_i P sigma^x_i P  -  Delta(t) * sum_i n_i,   |Omega| <= Rabi_max, |Delta| <= Delta_max
  channels: 0 = Omega/Rabi_max = 1 - exp(-x^2),  1 = Delta/Delta_max = tanh(x)      (no frequency filter)
  penalty : (1 - k/M) |g..g><g..g|  +  (k/M) |Z2><Z2|   in the top-right block of slice k
  cost    : 0.6 * (1 - |<Z2|U|g..g>|^2) + 0.4 * Re(1 - (U psi)^dag (W psi) / T),  psi = |g..g>
"""
from __future__ import annotations

import numpy as np

import qocvl
from .._engine import EngineBase, gradient_step, halving_flags


def blockade_basis(n_atoms: int):
    """Bit patterns (bit i = atom i excited) with no two adjacent excitations, ascending; F_{n+2} of them."""
    return [s for s in range(1 << n_atoms) if not (s & (s >> 1))]


def chain_generators(n_atoms: int, Rabi_max: float, Delta_max: float):
    """-> (states, gens_u [4,n,n], gens_w [4,n,n], index of |g..g>, index of |Z2>)."""
    states = blockade_basis(n_atoms)
    index = {s: k for k, s in enumerate(states)}
    n = len(states)
    drive = np.zeros((n, n), dtype=np.complex128)
    for s in states:
        for i in range(n_atoms):
            flipped = s ^ (1 << i)
            if flipped in index:                       # P sigma^x_i P: both patterns obey the blockade
                drive[index[flipped], index[s]] = 0.5 * Rabi_max
    detune = np.diag([-Delta_max * bin(s).count("1") for s in states]).astype(np.complex128)
    z2 = sum(1 << i for i in range(0, n_atoms, 2))     # 1010...: atoms 0, 2, 4, ... excited
    zero = np.zeros((n, n), dtype=np.complex128)
    proj_g, proj_z = zero.copy(), zero.copy()
    proj_g[index[0], index[0]] = 1.0
    proj_z[index[z2], index[z2]] = 1.0
    return states, np.stack([drive, detune, zero, zero]), np.stack([zero, zero, proj_g, proj_z]), index[0], index[z2]


class PropagatorVL(EngineBase):
    _model = qocvl.MODEL_CHAIN
    W_INFID, W_ADIAB = 0.6, 0.4
    num_ctrls = 2

    def __init__(self, input_dim: int, no_of_steps: int, delta_t: float, n_atoms: int, Rabi_max: float,
                 Delta_max: float, *, device=None):
        self.input_dim, self.no_of_steps, self.delta_t = input_dim, no_of_steps, delta_t
        self.n_atoms, self.Rabi_max, self.Delta_max = n_atoms, Rabi_max, Delta_max
        self.duration = no_of_steps * delta_t
        self.del_total = 0.0
        states, gens_u, gens_w, i_g, i_z = chain_generators(n_atoms, Rabi_max, Delta_max)
        self.states, self.dim = states, len(states)
        eye = np.eye(self.dim, dtype=np.complex128)
        self.psi_0, self.psi_t = eye[i_g], eye[i_z]
        self.contraction_array = halving_flags(no_of_steps)
        self.generators_u, self.generators_w = gens_u, gens_w
        self._make_plan(n=self.dim, n_slices=no_of_steps, n_basis_pts=no_of_steps, pad=0, gens_u=gens_u,
                        gens_w=gens_w, coef_const=np.zeros(4, complex), starts=[self.psi_0],
                        targets=[self.psi_t], adiab_index=0, fid_norm=1.0, scale=(Rabi_max, Delta_max, 0, 0),
                        taps=None, device=device)

    def _noise_tables(self, del_total, rabi_scale):
        """Sample s: Delta(t) -> Delta(t) + del_total[s], Omega(t) -> rabi_scale[s] * Omega(t)."""
        S = len(del_total)
        mul = np.ones((S, 4))
        mul[:, 0] = rabi_scale
        add = np.zeros((S, 4), complex)
        add[:, 1] = -1j * np.asarray(del_total) / self.Delta_max
        return mul, add


def single_optimization_step(propagator, ctrl_a, ctrl_b, ctrl_c, lr: float = 0.02):
    return gradient_step(propagator, ctrl_a, ctrl_b, ctrl_c, lr)
