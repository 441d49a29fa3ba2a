"""Robustness-ensemble oracle (SURVEY.md 8(d) cfg 3): the reference has no noise ensemble, so the
semantic is DEFINED here in the reference's own terms -- sample s is the reference class
re-instantiated with ``del_total = delta_s`` and both Rabi frequencies multiplied by
``(1 + eps_s)`` (TQ:49-62 constructor arguments); cost and gradient are plain means.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import numpy as np

from . import autograd_ref, two_qubit


def robust_cz_noise(n_samples, seed=1234, sigma_delta=2 * np.pi * 0.1e6, sigma_eps=0.01):
    """delta_s ~ N(0, 2pi*0.1 MHz), eps_s ~ N(0, 0.01): returns (del_total[S], rabi_scale[S])."""
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, sigma_delta, n_samples), 1.0 + rng.normal(0.0, sigma_eps, n_samples)


def robust_cz_members(del_total, rabi_scale, seed=0, n_gauss=16, nt=1888):
    """One oracle object per sample plus the shared (a, b, c)."""
    members, abc = [], None
    for d, s in zip(del_total, rabi_scale):
        prop, a, b, c = two_qubit.notebook02_setup(seed=seed, n_gauss=n_gauss, nt=nt, del_total=float(d),
                                                   rabi_scale=float(s))
        members.append(prop)
        abc = (a, b, c)
    return members, abc


def ensemble_value_and_grad(members, a, b, c):
    """Mean of autograd_ref.value_and_grad over the ensemble members."""
    acc = None
    for prop in members:
        res = autograd_ref.value_and_grad(prop, a, b, c)
        vals = [np.asarray(r, dtype=np.float64) for r in res]
        acc = vals if acc is None else [x + y for x, y in zip(acc, vals)]
    return tuple(x / len(members) for x in acc)
