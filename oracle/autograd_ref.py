"""torch-CPU complex128 autograd restatement: the gradient oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates ``jax.value_and_grad(target,
argnums=(0,1,2))`` (TQ:592, notebook 01 cell 9) by running the same forward chain as
oracle/state_transfer.py and oracle/two_qubit.py with torch ops and calling autograd.
``torch.linalg.matrix_exp`` replaces ``jax.scipy.linalg.expm`` (both backward-stable;
SURVEY.md 7.5-5).  Also the CPU baseline that bench.py times.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import state_transfer as st
from . import two_qubit as tq

F64, C128 = torch.float64, torch.complex128


def _modes(n_pts, centres, widths):
    t = torch.linspace(-1.0, 1.0, n_pts, dtype=F64)
    return torch.exp(-((t[:, None] - centres[None, :]) ** 2) / (widths[None, :] ** 2))


def _transform(n_pts, a, b, c):
    """ST:187-202 / TQ:335-353."""
    cb, cc = torch.tanh(b), torch.tanh(c)
    return torch.stack([_modes(n_pts, cb[:, j], cc[:, j]) @ a[:, j] for j in range(a.shape[1])], dim=1)


def _tree(ops, flags):
    """ST:258-283 / TQ:489-535."""
    for odd in flags:
        if odd:
            leftover, ops = ops[-1], ops[:-1]
            merged = torch.matmul(ops[1::2], ops[0:-1:2])
            ops = torch.cat([merged[:-1], (leftover @ merged[-1])[None]], dim=0)
        else:
            ops = torch.matmul(ops[1::2], ops[0::2])
    return ops.squeeze(0)


def wave_st(prop: st.PropagatorVL, a, b, c):
    """ST:215-227."""
    raw = _transform(prop.no_of_steps, a, b, c)

    def squash(pair):  # ST:204-213
        radius = torch.sqrt(pair[:, 0] ** 2 + pair[:, 1] ** 2)
        safe = torch.where(radius == 0, torch.ones_like(radius), radius)
        scale = torch.where(radius == 0, torch.ones_like(radius), torch.tanh(safe) / safe)
        return scale[:, None] * pair

    return torch.cat([torch.tanh(raw[:, 0:1]), squash(raw[:, 1:3]), squash(raw[:, 3:5])], dim=1)


def aux_st(prop: st.PropagatorVL, wave):
    """ST:236-245."""
    om1 = prop.Rabi_1 * torch.complex(wave[:, 1:2], wave[:, 2:3])
    om2 = prop.Rabi_2 * torch.complex(wave[:, 3:4], wave[:, 4:5])
    p1, p2 = om2 * om2.conj(), om1 * om1.conj()
    p3, p4 = -om1 * om2, -om1.conj() * om2.conj()
    weights = torch.cat([p1, p2, p3, p4], dim=1) / (p1 + p2)
    return torch.cat([-1j * wave.to(C128), weights], dim=1)


def wave_tq(prop: tq.PropagatorVL, a, b, c):
    """TQ:355-404."""
    raw = _transform(prop.no_of_steps, a, b, c)
    reg = torch.cat([0.0 * raw[:, 0:1] - 1.0,
                     1.0 - torch.exp(-raw[:, 1:2] ** 2), torch.tanh(raw[:, 2:3]),
                     1.0 - torch.exp(-raw[:, 3:4] ** 2), torch.tanh(raw[:, 4:5])], dim=1)
    pad = prop.padding
    padded = torch.nn.functional.pad(reg.to(C128), (0, 0, pad, pad))
    window = torch.from_numpy(prop.Gaussian_filter)
    spectrum = torch.fft.fftshift(torch.fft.fft(padded, dim=0), dim=0)
    back = torch.fft.ifft(torch.fft.ifftshift(window[:, None] * spectrum, dim=0), dim=0)
    return back.real


def aux_tq(prop: tq.PropagatorVL, wave):
    """TQ:416-455."""
    om_i = (prop.Rabi_i_val * wave[:, 1]) * torch.exp(1j * (math.pi * wave[:, 2]))
    om_r = (prop.Rabi_r_val * wave[:, 3]) * torch.exp(1j * (math.pi * wave[:, 4]))
    p1, p2 = om_r * om_r.conj(), om_i * om_i.conj()
    p3, p4 = -(om_i * om_r), -(om_i.conj() * om_r.conj())
    denom = p1 + p2
    denom = torch.where(denom.real < prop.EPS, torch.full_like(denom, prop.EPS), denom)
    weights = torch.stack([p1 / denom, p2 / denom, p3 / denom, p4 / denom], dim=1)
    ham = torch.stack([wave[:, 0], om_i.real / prop.Rabi_i_val, om_i.imag / prop.Rabi_i_val,
                       om_r.real / prop.Rabi_r_val, om_r.imag / prop.Rabi_r_val], dim=1).to(C128)
    return torch.cat([-1j * ham, weights], dim=1)


def target_torch(prop, a, b, c, return_parts=False):
    """Differentiable cost for either oracle class (dispatch on type)."""
    gens = torch.from_numpy(np.ascontiguousarray(prop.VL_Generators_ansatz))
    if isinstance(prop, tq.PropagatorVL):
        aux = aux_tq(prop, wave_tq(prop, a, b, c))
        drift = -1j * (gens[0] + gens[1])                       # TQ:470
    else:
        aux = aux_st(prop, wave_st(prop, a, b, c))
        drift = gens[0] + gens[1]                               # ST:254
    exponent = prop.delta_t * (drift[None] + torch.tensordot(aux, gens[2:], dims=([1], [0])))
    vl = _tree(torch.linalg.matrix_exp(exponent), prop.contraction_array)
    if isinstance(prop, tq.PropagatorVL):
        n = 16
        u = vl[:n, :n]
        ut = torch.from_numpy(prop.U_target)
        tr = torch.trace(ut.conj().T @ u)
        infid = 1.0 - ((tr * tr.conj()) / 16.0).real            # (Tr Ut^dag Ut)^2 = 16
        psi = torch.from_numpy(prop.psi_10)
    else:
        n = 5
        u = vl[:n, :n]
        psi = torch.from_numpy(prop.psi_0)
        overlap = torch.from_numpy(prop.psi_t).conj() @ (u @ psi)
        infid = 1.0 - (overlap.conj() * overlap).real
    final = u @ psi
    adiab = (1.0 - (1.0 / prop.duration) * (final.conj() @ (vl[:n, n:2 * n] @ psi))).real
    cost = prop.W_INFID * infid + prop.W_ADIAB * adiab
    return (cost, infid, adiab) if return_parts else cost


def value_and_grad(prop, a, b, c):
    """-> (cost, infid, adiab, ga, gb, gc) as floats / float64 numpy arrays shaped like a,b,c."""
    ta, tb, tc = (torch.tensor(np.asarray(x, dtype=np.float64), requires_grad=True) for x in (a, b, c))
    cost, infid, adiab = target_torch(prop, ta, tb, tc, return_parts=True)
    ga, gb, gc = torch.autograd.grad(cost, (ta, tb, tc))
    return (float(cost.detach()), float(infid.detach()), float(adiab.detach()), ga.numpy(), gb.numpy(), gc.numpy())
