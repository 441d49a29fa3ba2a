"""Oracle for the blockade-chain (PXP) extrapolation of BASELINE.json configs 4-5.
TEST INFRASTRUCTURE (see oracle/__init__.py).  There is no reference code for this model: the oracle
DEFINES it by running the reference's own pipeline shape -- Gaussian basis (ST:173-202), regularisation,
dense Van Loan generators [[H,P],[0,H]] (ST:132-170), expm per slice, ordered product (ST:258-283), the
ST cost (ST:285-310) -- with torch autograd for the gradient.  Dense 2n x 2n matrices, independent of the
sparse CUDA path."""
from __future__ import annotations

import numpy as np
import torch

from .state_transfer import halving_flags

C128 = torch.complex128


def basis_states(n_atoms):
    out = []
    for s in range(2 ** n_atoms):
        bits = [(s >> i) & 1 for i in range(n_atoms)]
        if all(not (bits[i] and bits[i + 1]) for i in range(n_atoms - 1)):
            out.append(s)
    return out


class BlockadeChain:
    W_INFID, W_ADIAB = 0.6, 0.4

    def __init__(self, input_dim, no_of_steps, delta_t, n_atoms, Rabi_max, Delta_max, del_offset=0.0, rabi_scale=1.0):
        self.input_dim, self.no_of_steps, self.delta_t = input_dim, no_of_steps, delta_t
        self.duration = no_of_steps * delta_t
        self.n_atoms, self.Rabi_max, self.Delta_max = n_atoms, Rabi_max, Delta_max
        self.del_offset, self.rabi_scale = del_offset, rabi_scale
        st = basis_states(n_atoms)
        pos = {s: i for i, s in enumerate(st)}
        n = self.dim = len(st)
        hx = np.zeros((n, n), complex)
        for s in st:
            for i in range(n_atoms):
                t = s ^ (1 << i)
                if t in pos:
                    hx[pos[t], pos[s]] += 0.5
        nn = np.diag([bin(s).count("1") for s in st]).astype(complex)
        self.h_drive, self.h_number = hx, nn
        self.i_g = pos[0]
        self.i_z = pos[sum(1 << i for i in range(0, n_atoms, 2))]
        self.contraction_array = halving_flags(no_of_steps)

    def waves(self, a, b, c):
        """[M,2]: Omega/Rabi_max = 1 - exp(-x^2), Delta/Delta_max = tanh(x); torch in, torch out."""
        t = torch.linspace(-1.0, 1.0, self.no_of_steps, dtype=torch.float64)
        bt, ct = torch.tanh(b), torch.tanh(c)
        raw = torch.stack([torch.exp(-((t[:, None] - bt[None, :, j]) ** 2) / ct[None, :, j] ** 2) @ a[:, j]
                           for j in range(2)], dim=1)
        return torch.stack([1.0 - torch.exp(-raw[:, 0] ** 2), torch.tanh(raw[:, 1])], dim=1)

    def cost_torch(self, a, b, c):
        n, M = self.dim, self.no_of_steps
        w = self.waves(a, b, c)
        hx, nn = torch.from_numpy(self.h_drive), torch.from_numpy(self.h_number)
        omega = self.rabi_scale * self.Rabi_max * w[:, 0]
        delta = self.Delta_max * w[:, 1] + self.del_offset
        ham = omega[:, None, None] * hx[None] - delta[:, None, None] * nn[None]
        ramp = torch.arange(M, dtype=torch.float64) / M
        pen = torch.zeros((M, n, n), dtype=C128)
        pen[:, self.i_g, self.i_g] = (1.0 - ramp).to(C128)
        pen[:, self.i_z, self.i_z] = ramp.to(C128)
        top = torch.cat([-1j * ham, pen], dim=2)
        bot = torch.cat([torch.zeros_like(ham), -1j * ham], dim=2)
        steps = torch.linalg.matrix_exp(self.delta_t * torch.cat([top, bot], dim=1))
        for odd in self.contraction_array:                      # ST:258-283
            if odd:
                last, steps = steps[-1], steps[:-1]
                merged = torch.matmul(steps[1::2], steps[0:-1:2])
                steps = torch.cat([merged[:-1], (last @ merged[-1])[None]], dim=0)
            else:
                steps = torch.matmul(steps[1::2], steps[0::2])
        vl = steps.squeeze(0) if steps.dim() == 3 else steps
        u, wblk = vl[:n, :n], vl[:n, n:]
        final = u[:, self.i_g]
        infid = 1.0 - (final[self.i_z].conj() * final[self.i_z]).real
        adiab = (1.0 - (final.conj() @ wblk[:, self.i_g]) / self.duration).real
        return self.W_INFID * infid + self.W_ADIAB * adiab, infid, adiab

    def value_and_grad(self, a, b, c):
        ta, tb, tc = (torch.tensor(np.asarray(x, float), requires_grad=True) for x in (a, b, c))
        cost, infid, adiab = self.cost_torch(ta, tb, tc)
        ga, gb, gc = torch.autograd.grad(cost, (ta, tb, tc))
        return float(cost.detach()), float(infid.detach()), float(adiab.detach()), ga.numpy(), gb.numpy(), gc.numpy()


def chain_setup(n_atoms=5, n_gauss=6, nt=64, seed=7, total_time=None, del_offset=0.0, rabi_scale=1.0):
    """SURVEY.md 8(d) cfg 4 parameters at reduced size: Omega_max = 2pi*2 MHz, Delta_max = 2pi*4 MHz, dt = 1 ns."""
    dt = 1e-9 if total_time is None else total_time / nt
    # reduced grids keep a comparable |X_k| by using a coarser step (20 ns) so that the pulses actually act
    if total_time is None and nt < 1000:
        dt = 20e-9
    prop = BlockadeChain(n_gauss, nt, dt, n_atoms, 2 * np.pi * 2e6, 2 * np.pi * 4e6, del_offset, rabi_scale)
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (n_gauss, 2)); b = rng.uniform(-1, 1, (n_gauss, 2)); c = rng.uniform(0, 0.5, (n_gauss, 2))
    return prop, a, b, c
