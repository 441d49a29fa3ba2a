"""State-vector restatement of the cost AND its gradient (manual adjoint), NumPy only.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the third, independent
formulation used to pin the oracle: instead of forming the Van Loan propagator
(ST:247-283, TQ:460-535) it pushes only the columns the cost reads (ST:294-301,
TQ:550-565) through the slices with a truncated Taylor series of exp(X_k), and
obtains the gradient from the exact reverse sweep of that recurrence followed by
hand-derived chain rules through the auxiliary amplitudes (ST:229-245, TQ:406-455),
the frequency filter (TQ:375-397), the regularisation (ST:204-227, TQ:355-373)
and the Gaussian basis (ST:173-202, TQ:319-353).

It is written with dense n x n matrices and plain Python loops (it is a checker, not
a fast path) but the recurrences are the ones the CUDA chain kernel implements, so
a failure localises to a stage.
"""
from __future__ import annotations

import math

import numpy as np

from . import state_transfer as st
from . import two_qubit as tq


# ----------------------------------------------------------------------------
# generator blocks and coefficient maps
# ----------------------------------------------------------------------------
def split_generators(gens):
    """[J,2n,2n] -> (u [J,n,n] top-left blocks, w [J,n,n] top-right blocks)."""
    n = gens.shape[1] // 2
    assert np.allclose(gens[:, n:, :n], 0) and np.allclose(gens[:, :n, :n], gens[:, n:, n:])
    return gens[:, :n, :n].copy(), gens[:, :n, n:].copy()


def coefficients_tq(prop: tq.PropagatorVL, wave):
    """coef [M,11] and d coef / d wave [M,11,5] for TQ:416-476 (drift gens get -1j, TQ:470)."""
    m = wave.shape[0]
    ri, rr = prop.Rabi_i_val, prop.Rabi_r_val
    w0, w1, w2, w3, w4 = (wave[:, k] for k in range(5))
    ci, si = np.cos(np.pi * w2), np.sin(np.pi * w2)
    cr, sr = np.cos(np.pi * w4), np.sin(np.pi * w4)
    coef = np.zeros((m, 11), complex)
    dco = np.zeros((m, 11, 5), complex)
    coef[:, 0] = coef[:, 1] = -1j
    coef[:, 2] = -1j * w0;          dco[:, 2, 0] = -1j
    coef[:, 3] = -1j * w1 * ci;     dco[:, 3, 1] = -1j * ci; dco[:, 3, 2] = 1j * np.pi * w1 * si
    coef[:, 4] = -1j * w1 * si;     dco[:, 4, 1] = -1j * si; dco[:, 4, 2] = -1j * np.pi * w1 * ci
    coef[:, 5] = -1j * w3 * cr;     dco[:, 5, 3] = -1j * cr; dco[:, 5, 4] = 1j * np.pi * w3 * sr
    coef[:, 6] = -1j * w3 * sr;     dco[:, 6, 3] = -1j * sr; dco[:, 6, 4] = -1j * np.pi * w3 * cr
    mi, mr = ri * w1, rr * w3
    den_raw = mr * mr + mi * mi
    clamped = den_raw < prop.EPS
    den = np.where(clamped, prop.EPS, den_raw)
    dden1 = np.where(clamped, 0.0, 2 * mi * ri)
    dden3 = np.where(clamped, 0.0, 2 * mr * rr)
    ph = np.exp(1j * np.pi * (w2 + w4))
    coef[:, 7] = mr * mr / den
    coef[:, 8] = mi * mi / den
    coef[:, 9] = -mi * mr * ph / den
    coef[:, 10] = np.conj(coef[:, 9])
    dco[:, 7, 1] = -mr * mr * dden1 / den ** 2
    dco[:, 7, 3] = 2 * mr * rr / den - mr * mr * dden3 / den ** 2
    dco[:, 8, 1] = 2 * mi * ri / den - mi * mi * dden1 / den ** 2
    dco[:, 8, 3] = -mi * mi * dden3 / den ** 2
    dco[:, 9, 1] = -ph * (ri * mr / den - mi * mr * dden1 / den ** 2)
    dco[:, 9, 3] = -ph * (mi * rr / den - mi * mr * dden3 / den ** 2)
    dco[:, 9, 2] = 1j * np.pi * coef[:, 9]
    dco[:, 9, 4] = 1j * np.pi * coef[:, 9]
    dco[:, 10, :] = np.conj(dco[:, 9, :])
    return coef, dco


def coefficients_st(prop: st.PropagatorVL, wave):
    """coef [M,11], dcoef [M,11,5] for ST:236-255 (drift gens get +1: quirk q1)."""
    m = wave.shape[0]
    r1, r2 = prop.Rabi_1, prop.Rabi_2
    coef = np.zeros((m, 11), complex)
    dco = np.zeros((m, 11, 5), complex)
    coef[:, 0] = coef[:, 1] = 1.0
    for ch in range(5):
        coef[:, 2 + ch] = -1j * wave[:, ch]
        dco[:, 2 + ch, ch] = -1j
    o1 = r1 * (wave[:, 1] + 1j * wave[:, 2])
    o2 = r2 * (wave[:, 3] + 1j * wave[:, 4])
    n1, n2 = np.abs(o1) ** 2, np.abs(o2) ** 2
    with np.errstate(invalid="ignore", divide="ignore"):
        den = n1 + n2
        coef[:, 7] = n2 / den
        coef[:, 8] = n1 / den
        coef[:, 9] = -o1 * o2 / den
        coef[:, 10] = np.conj(coef[:, 9])
        dn1 = [None, 2 * r1 * r1 * wave[:, 1], 2 * r1 * r1 * wave[:, 2], 0 * den, 0 * den]
        dn2 = [None, 0 * den, 0 * den, 2 * r2 * r2 * wave[:, 3], 2 * r2 * r2 * wave[:, 4]]
        do1 = [None, r1 + 0j, 1j * r1, 0j, 0j]
        do2 = [None, 0j, 0j, r2 + 0j, 1j * r2]
        for ch in range(1, 5):
            dden = dn1[ch] + dn2[ch]
            dco[:, 7, ch] = dn2[ch] / den - n2 * dden / den ** 2
            dco[:, 8, ch] = dn1[ch] / den - n1 * dden / den ** 2
            dco[:, 9, ch] = -(do1[ch] * o2 + o1 * do2[ch]) / den + o1 * o2 * dden / den ** 2
        dco[:, 10, :] = np.conj(dco[:, 9, :])
    return coef, dco


# ----------------------------------------------------------------------------
# pulse map with hand-written VJP
# ----------------------------------------------------------------------------
def filter_taps(prop: tq.PropagatorVL):
    """Real impulse response h[M] of TQ:375-397: wave[k] = sum_t h[(k - t - pad) mod M] reg[t]."""
    m = prop.no_of_steps + 2 * prop.padding
    impulse = np.zeros(m, complex)
    impulse[0] = 1.0
    spec = np.fft.fftshift(np.fft.fft(impulse))
    return np.real(np.fft.ifft(np.fft.ifftshift(prop.Gaussian_filter * spec)))


def pulse_map(prop, a, b, c, model):
    """-> wave [M,5] and a cache for pulse_map_vjp.  model in {"tq","st"}."""
    nt = prop.no_of_steps
    t = np.linspace(-1.0, 1.0, nt)
    bt, ct = np.tanh(b), np.tanh(c)
    modes = np.exp(-((t[:, None, None] - bt[None]) ** 2) / (ct[None] ** 2))   # [nt,N,5]
    raw = np.einsum("tnc,nc->tc", modes, a)
    if model == "tq":
        reg = np.stack([-np.ones(nt), 1 - np.exp(-raw[:, 1] ** 2), np.tanh(raw[:, 2]),
                        1 - np.exp(-raw[:, 3] ** 2), np.tanh(raw[:, 4])], axis=1)
        taps = filter_taps(prop)
        m, pad = nt + 2 * prop.padding, prop.padding
        idx = (np.arange(m)[:, None] - np.arange(nt)[None, :] - pad) % m
        kmat = taps[idx]                                                      # [M,nt]
        wave = kmat @ reg
    else:
        reg = np.empty_like(raw)
        reg[:, 0] = np.tanh(raw[:, 0])
        for p0 in (1, 3):
            rad = np.hypot(raw[:, p0], raw[:, p0 + 1])
            with np.errstate(invalid="ignore", divide="ignore"):
                scale = np.where(rad == 0, 1.0, np.tanh(rad) / rad)
            reg[:, p0] = scale * raw[:, p0]
            reg[:, p0 + 1] = scale * raw[:, p0 + 1]
        kmat, wave = None, reg
    return wave, dict(t=t, bt=bt, ct=ct, modes=modes, raw=raw, kmat=kmat, a=a, model=model)


def pulse_map_vjp(cache, gwave):
    """d cost / d (a,b,c) from d cost / d wave [M,5]."""
    raw, model = cache["raw"], cache["model"]
    greg = cache["kmat"].T @ gwave if model == "tq" else gwave
    graw = np.zeros_like(raw)
    if model == "tq":
        for ch in (1, 3):
            graw[:, ch] = greg[:, ch] * 2 * raw[:, ch] * np.exp(-raw[:, ch] ** 2)
        for ch in (2, 4):
            graw[:, ch] = greg[:, ch] * (1 - np.tanh(raw[:, ch]) ** 2)
    else:
        graw[:, 0] = greg[:, 0] * (1 - np.tanh(raw[:, 0]) ** 2)
        for p0 in (1, 3):
            x, y = raw[:, p0], raw[:, p0 + 1]
            rad = np.hypot(x, y)
            with np.errstate(invalid="ignore", divide="ignore"):
                th = np.tanh(rad)
                scale = np.where(rad == 0, 1.0, th / rad)
                # d scale / d rad / rad  (0 where rad == 0: jnp.where picks the constant branch)
                dsc = np.where(rad == 0, 0.0, ((1 - th * th) * rad - th) / rad ** 3)
            dot = greg[:, p0] * x + greg[:, p0 + 1] * y
            graw[:, p0] = scale * greg[:, p0] + dsc * dot * x
            graw[:, p0 + 1] = scale * greg[:, p0 + 1] + dsc * dot * y
    t, bt, ct, modes, a = cache["t"], cache["bt"], cache["ct"], cache["modes"], cache["a"]
    dev = t[:, None, None] - bt[None]
    ga = np.einsum("tnc,tc->nc", modes, graw)
    gb = a * np.einsum("tnc,tc->nc", modes * 2 * dev / ct[None] ** 2, graw) * (1 - bt ** 2)
    gc = a * np.einsum("tnc,tc->nc", modes * 2 * dev ** 2 / ct[None] ** 3, graw) * (1 - ct ** 2)
    return ga, gb, gc


# ----------------------------------------------------------------------------
# Taylor degree rule shared with the CUDA host code (qocvl/csrc/plan.cpp: taylor_degree)
# ----------------------------------------------------------------------------
def taylor_degree(norm_bound, tol=2.0 ** -60, q_max=40):
    """Smallest q with norm^(q+1)/(q+1)! <= tol (remainder bound of the truncated series)."""
    term, q = 1.0, 0
    while q < q_max:
        q += 1
        term *= norm_bound / q
        if term * norm_bound / (q + 1) <= tol and q >= 2:
            return q
    return q_max


# ----------------------------------------------------------------------------
# the chain: forward columns, reverse sweep
# ----------------------------------------------------------------------------
class CostSpec:
    """What the cost reads: start kets x_v, target kets y_v, the adiabatic ket index."""

    def __init__(self, starts, targets, adiab_index, norm, w_infid, w_adiab, duration):
        self.starts, self.targets = np.asarray(starts, complex), np.asarray(targets, complex)
        self.adiab_index, self.norm = adiab_index, norm
        self.w_infid, self.w_adiab, self.duration = w_infid, w_adiab, duration


def cost_spec_for(prop):
    if isinstance(prop, tq.PropagatorVL):
        cols = [0, 1, 4, 5]                      # non-zero columns of U_target (TQ:121-127)
        eye = np.eye(16, dtype=complex)
        return CostSpec([eye[c] for c in cols], [prop.U_target[:, c] for c in cols],
                        cols.index(4), 16.0, prop.W_INFID, prop.W_ADIAB, prop.duration)
    return CostSpec([prop.psi_0], [prop.psi_t], 0, 1.0, prop.W_INFID, prop.W_ADIAB, prop.duration)


def chain_cost_grad(gu, gw, coef, dcoef, dt, spec: CostSpec, want_grad=True, tol=2.0 ** -60):
    """cost, infid, adiab and d cost / d wave [M,C] for X_k = dt * sum_j coef[k,j] (gu_j, gw_j)."""
    m, n = coef.shape[0], gu.shape[1]
    n_vec = len(spec.starts)
    va = spec.adiab_index
    z = [spec.starts[v].copy() for v in range(n_vec)]    # U x_v
    p = np.zeros(n, complex)                             # W psi
    saved, degs = [], []
    norms_u = np.abs(gu).sum(axis=1).max(axis=1)
    norms_w = np.abs(gw).sum(axis=1).max(axis=1)
    for k in range(m):
        a_mat = dt * np.tensordot(coef[k], gu, axes=[[0], [0]])
        b_mat = dt * np.tensordot(coef[k], gw, axes=[[0], [0]])
        q = taylor_degree(dt * float(np.abs(coef[k]) @ (norms_u + norms_w)), tol)
        degs.append(q)
        saved.append(([zz.copy() for zz in z], p.copy()))
        tz, tp = [zz.copy() for zz in z], p.copy()
        for j in range(1, q + 1):
            tp = (a_mat @ tp + b_mat @ tz[va]) / j
            tz = [a_mat @ x / j for x in tz]
            p = p + tp
            z = [zz + x for zz, x in zip(z, tz)]
    d = sum(np.vdot(spec.targets[v], z[v]) for v in range(n_vec))
    infid = 1.0 - abs(d) ** 2 / spec.norm
    adiab = 1.0 - np.real(np.vdot(z[va], p)) / spec.duration
    cost = spec.w_infid * infid + spec.w_adiab * adiab
    if not want_grad:
        return cost, infid, adiab, None
    # terminal cotangents  lambda = d cost / d conj(z)   (d cost = 2 Re sum lambda^dag dz)
    lz = [-(spec.w_infid / spec.norm) * d * spec.targets[v] for v in range(n_vec)]
    lz[va] = lz[va] - (spec.w_adiab / (2 * spec.duration)) * p
    lp = -(spec.w_adiab / (2 * spec.duration)) * z[va]
    gwave = np.zeros((m, dcoef.shape[2]))
    n_gen = gu.shape[0]
    for k in range(m - 1, -1, -1):
        a_mat = dt * np.tensordot(coef[k], gu, axes=[[0], [0]])
        b_mat = dt * np.tensordot(coef[k], gw, axes=[[0], [0]])
        q = degs[k]
        z0, p0 = saved[k]
        # forward Taylor terms t_0 .. t_{q-1}
        tz, tp = [[x.copy() for x in z0]], [p0.copy()]
        for j in range(1, q):
            tp.append((a_mat @ tp[-1] + b_mat @ tz[-1][va]) / j)
            tz.append([a_mat @ x / j for x in tz[-1]])
        # reverse: tbar_q = lambda'; tbar_{j-1} = lambda' + X^dag tbar_j / j
        bz, bp = [x.copy() for x in lz], lp.copy()
        xbar_a = np.zeros((n, n), complex)      # sum_j conj(tbar_j[a]) t_{j-1}[b] / j
        xbar_b = np.zeros((n, n), complex)
        ah, bh = a_mat.conj().T, b_mat.conj().T
        for j in range(q, 0, -1):
            for v in range(n_vec):
                xbar_a += np.outer(bz[v].conj(), tz[j - 1][v]) / j
            xbar_a += np.outer(bp.conj(), tp[j - 1]) / j
            xbar_b += np.outer(bp.conj(), tz[j - 1][va]) / j
            new_bz = [lz[v] + ah @ bz[v] / j for v in range(n_vec)]
            new_bz[va] = new_bz[va] + bh @ bp / j
            bp = lp + ah @ bp / j
            bz = new_bz
        lz, lp = bz, bp
        sens = np.array([2 * dt * (np.sum(gu[g] * xbar_a) + np.sum(gw[g] * xbar_b))
                         for g in range(n_gen)])                 # d cost = Re sum sens_j dcoef_j
        gwave[k] = np.real(sens @ dcoef[k])
    return cost, infid, adiab, gwave


def value_and_grad(prop, a, b, c):
    """Same return convention as oracle.autograd_ref.value_and_grad."""
    a, b, c = (np.asarray(x, float) for x in (a, b, c))
    model = "tq" if isinstance(prop, tq.PropagatorVL) else "st"
    wave, cache = pulse_map(prop, a, b, c, model)
    coef, dcoef = (coefficients_tq if model == "tq" else coefficients_st)(prop, wave)
    gu, gw = split_generators(prop.VL_Generators_ansatz)
    cost, infid, adiab, gwave = chain_cost_grad(gu, gw, coef, dcoef, prop.delta_t, cost_spec_for(prop))
    ga, gb, gc = pulse_map_vjp(cache, gwave)
    return cost, infid, adiab, ga, gb, gc
