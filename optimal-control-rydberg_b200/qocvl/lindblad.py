"""Host wrapper of ``qocvl_lindblad_propagate`` (csrc/lindblad.cu): vectorised-density-matrix propagation under a
time-dependent sparse Liouvillian, exponential midpoint rule.  NumPy in / NumPy out (a one-shot validation call)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def union_csr(mats, tol=0.0):
    """[J] dense (n, n) complex matrices -> (rowptr[n+1], col[nnz], vals[J, nnz]) on their union pattern."""
    mats = [np.asarray(m, dtype=np.complex128) for m in mats]
    n = mats[0].shape[0]
    pattern = np.zeros((n, n), dtype=bool)
    for m in mats:
        pattern |= np.abs(m) > tol
    rows, cols = np.nonzero(pattern)                 # row-major order: CSR
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    vals = np.stack([m[rows, cols] for m in mats]).astype(np.complex128)
    return rowptr, cols.astype(np.int32), vals


def propagate(mats, coef, rho0, dt, obs=None, device=None):
    """rho_{k+1} = exp(dt * sum_j coef[..., k, j] * mats[j]) rho_k for k = 0 .. n_steps-1.

    mats: [J] dense (n, n) Liouvillian pieces; coef: [n_steps, J] or [n_traj, n_steps, J] (complex);
    rho0: [n] or [n_traj, n] vectorised initial states; obs: optional [n_obs, n] row vectors w (expectation =
    w . rho).  Returns (expect [n_traj, n_steps+1, n_obs] complex, rho_final [n_traj, n] complex)."""
    if not torch.cuda.is_available():
        raise RuntimeError("qocvl.lindblad needs a CUDA device (there is no CPU fallback)")
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    rowptr, col, vals = union_csr(mats)
    n, nnz, J = rowptr.shape[0] - 1, col.shape[0], vals.shape[0]
    coef = np.ascontiguousarray(np.asarray(coef, dtype=np.complex128))
    rho0 = np.ascontiguousarray(np.asarray(rho0, dtype=np.complex128))
    if coef.ndim == 2:
        coef = coef[None]
    if rho0.ndim == 1:
        rho0 = rho0[None]
    n_traj = max(coef.shape[0], rho0.shape[0])
    if coef.shape[0] == 1 and n_traj > 1:
        coef = np.ascontiguousarray(np.broadcast_to(coef, (n_traj,) + coef.shape[1:]))
    if rho0.shape[0] == 1 and n_traj > 1:
        rho0 = np.ascontiguousarray(np.broadcast_to(rho0, (n_traj, n)))
    if coef.shape[0] != n_traj or rho0.shape != (n_traj, n) or coef.shape[2] != J:
        raise ValueError(f"shape mismatch: coef {coef.shape}, rho0 {rho0.shape}, {J} generators of size {n}")
    n_steps = coef.shape[1]
    obs = np.zeros((0, n), complex) if obs is None else np.ascontiguousarray(np.asarray(obs, dtype=np.complex128))
    n_obs = obs.shape[0]
    if n_obs and obs.shape != (n_obs, n):
        raise ValueError("obs must be [n_obs, n]")
    expect = np.zeros((n_traj, n_steps + 1, max(n_obs, 1)), dtype=np.complex128)
    rho_out = np.zeros((n_traj, n), dtype=np.complex128)
    vals = np.ascontiguousarray(vals)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    st = lib.qocvl_lindblad_propagate(dev.index or 0, n, nnz, J, n_steps, n_traj, n_obs, ptr(rowptr), ptr(col),
                                      ptr(vals), ptr(coef), ptr(rho0), ptr(obs) if n_obs else None, float(dt),
                                      ptr(expect) if n_obs else None, ptr(rho_out))
    _lib.check(st, "lindblad_propagate")
    return expect[:, :, :n_obs], rho_out
