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

_B13 = (64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
        10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.)


def expm_pade13(x):
    """Batched, differentiable matrix exponential: Higham's (2005) scaling-and-squaring Pade-13 approximant in torch ops
    (what jax.scipy.linalg.expm -- the reference's expm, ST:256 / TQ:480 -- implements for its largest norm class).
    Used instead of torch.linalg.matrix_exp because the latter loses up to 2e-10 per slice on these generators when the
    1-norm falls just under its degree-8 threshold (0.0499; measured against scipy.linalg.expm while building the
    full-size fixtures), which is not good enough for a 1e-10 parity bound over thousands of slices."""
    norm = float(x.detach().abs().sum(dim=-2).max())
    s = max(0, int(np.ceil(np.log2(norm / 5.371920351148152)))) if norm > 5.371920351148152 else 0
    a = x / (2.0 ** s)
    ident = torch.eye(x.shape[-1], dtype=x.dtype).expand_as(a)
    a2 = a @ a
    a4 = a2 @ a2
    a6 = a4 @ a2
    b = _B13
    u = a @ (a6 @ (b[13] * a6 + b[11] * a4 + b[9] * a2) + b[7] * a6 + b[5] * a4 + b[3] * a2 + b[1] * ident)
    v = a6 @ (b[12] * a6 + b[10] * a4 + b[8] * a2) + b[6] * a6 + b[4] * a4 + b[2] * a2 + b[0] * ident
    r = torch.linalg.solve(v - u, v + u)
    for _ in range(s):
        r = r @ r
    return r


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
        steps = expm_pade13(self.delta_t * torch.cat([top, bot], dim=1))
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

    def _dense_steps(self, w, lo, hi):
        """Dense Van Loan exponentials of slices lo..hi-1 from the waves w[M,2] (same construction as cost_torch)."""
        n, M = self.dim, self.no_of_steps
        hx, nn = torch.from_numpy(self.h_drive), torch.from_numpy(self.h_number)
        omega = self.rabi_scale * self.Rabi_max * w[lo:hi, 0]
        delta = self.Delta_max * w[lo:hi, 1] + self.del_offset
        ham = omega[:, None, None] * hx[None] - delta[:, None, None] * nn[None]
        ramp = torch.arange(lo, hi, dtype=torch.float64) / M
        pen = torch.zeros((hi - lo, n, n), dtype=C128)
        pen[:, self.i_g, self.i_g] = (1.0 - ramp).to(C128)
        pen[:, self.i_z, self.i_z] = ramp.to(C128)
        top = torch.cat([-1j * ham, pen], dim=2)
        bot = torch.cat([torch.zeros_like(ham), -1j * ham], dim=2)
        return expm_pade13(self.delta_t * torch.cat([top, bot], dim=1))

    def cost_torch_columns(self, a, b, c, chunk=125):
        """Memory-light twin of cost_torch for the FULL-SIZE configs (n = 144 x M = 4000; n = 377 x M = 256): the same
        dense 2n x 2n exponentials per slice, but instead of the product tree (which keeps ~2M dense matrices alive for
        autograd) the two columns of the Van Loan propagator the cost reads -- VL e_g = [U e_g; 0] and
        VL e_{n+g} = [W e_g; U e_g] -- are pushed through the slices in time order, chunk by chunk under
        torch.utils.checkpoint.  Same numbers up to the association order of the products (3e-16)."""
        from torch.utils.checkpoint import checkpoint
        n, M = self.dim, self.no_of_steps
        w = self.waves(a, b, c)
        v = torch.zeros((2 * n, 2), dtype=C128)
        v[self.i_g, 0] = 1.0
        v[n + self.i_g, 1] = 1.0

        def run(vec, waves, lo, hi):
            steps = self._dense_steps(waves, lo, hi)
            for k in range(hi - lo):
                vec = steps[k] @ vec
            return vec
        for lo in range(0, M, chunk):
            v = checkpoint(run, v, w, lo, min(M, lo + chunk), use_reentrant=False)
        final, wcol = v[:n, 0], v[:n, 1]
        infid = 1.0 - (final[self.i_z].conj() * final[self.i_z]).real
        adiab = (1.0 - (final.conj() @ wcol) / self.duration).real
        return self.W_INFID * infid + self.W_ADIAB * adiab, infid, adiab

    def value_and_grad_columns(self, a, b, c, chunk=125):
        ta, tb, tc = (torch.tensor(np.asarray(x, float), requires_grad=True) for x in (a, b, c))
        cost, infid, adiab = self.cost_torch_columns(ta, tb, tc, chunk)
        ga, gb, gc = torch.autograd.grad(cost, (ta, tb, tc))
        return float(cost.detach()), float(infid.detach()), float(adiab.detach()), ga.numpy(), gb.numpy(), gc.numpy()

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
