"""NumPy restatement of the engine's norm-aware Taylor-degree rule (CUDA: pulse_map.cu::k_slice_degrees,
chain_wide.cu::k_wide_degrees).  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference has no counterpart: it calls ``jax.scipy.linalg.expm`` (TQ:460-484, ST:247-256), which picks a Pade order
from the 1-norm of every slice matrix internally.  The engine pushes vectors through a truncated Taylor series instead
and needs, per slice and per independent diagonal block ("role") of the propagated system, a bound on the norm of the
block over every noise sample.  Three rigorous ingredients:

* ``collatz_wielandt_two_norm``:  |Y|_2 <= | |Y| |_2 <= |N|_2 = sqrt(rho(N^T N)) <= sqrt(max_i (N^T N v)_i / v_i) for ANY
  positive vector v, where N >= |Y| entrywise (entrywise absolute values only raise the 2-norm; it is monotone on
  non-negative matrices; the last step is the Collatz-Wielandt bound).  A few power steps make v the Perron vector.
* ``phase_shift``:  exp(X) = e^{i phi} exp(X - i phi I) for every real phi; phi = midpoint of the range of Im diag X.
* ``taylor_degree``: smallest q with nrm^(q+1)/(q+1)! <= tol (oracle/vector_model.py holds the same rule).
"""
from __future__ import annotations

import numpy as np

from .vector_model import taylor_degree


def collatz_wielandt_two_norm(n_mat, steps=5, floor=1e-3):
    """Upper bound of the 2-norm of the NON-NEGATIVE matrix ``n_mat`` (rigorous for any ``steps``)."""
    n_mat = np.asarray(n_mat, float)
    assert (n_mat >= 0).all()
    v = np.ones(n_mat.shape[1])
    for _ in range(steps):
        t = n_mat.T @ (n_mat @ v)
        mx = t.max()
        if mx <= 0:
            return 0.0
        v = np.maximum(t / mx, floor)
    t = n_mat.T @ (n_mat @ v)
    return float(np.sqrt((t / v).max()))


def phase_shift(x_mat):
    """phi = midpoint of the range of Im diag(X): the sweeps propagate with X - i phi I."""
    d = np.diag(np.asarray(x_mat)).imag
    return 0.5 * (d.max() + d.min())


def block_degree(x_nominal, deviation=None, tol=2.0 ** -55, shift=True, steps=5):
    """Taylor degree of one block: min(1-norm, Collatz-Wielandt 2-norm bound) of N = |X - i phi I| + deviation."""
    x_nominal = np.asarray(x_nominal, complex)
    phi = phase_shift(x_nominal) if shift else 0.0
    n_mat = np.abs(x_nominal - 1j * phi * np.eye(x_nominal.shape[0]))
    if deviation is not None:
        n_mat = n_mat + np.asarray(deviation, float)
    one = n_mat.sum(axis=0).max()
    two = collatz_wielandt_two_norm(n_mat, steps)
    return taylor_degree(min(one, two), tol), phi
