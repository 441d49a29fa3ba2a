"""Property tests (hypothesis) of the host logic and of the oracle's building blocks -- CPU only.

* `shard_bounds` is a partition of the samples into contiguous, balanced blocks for any (samples, world);
* the reference's halving product tree (ST:258-283 / TQ:489-535: later slice on the left, odd leftover folded into the
  tail) equals the plain time-ordered product for ANY number of slices and non-commuting factors;
* the engine's real filter taps reproduce the reference's zero-pad / fft / window / ifft / real-part chain (TQ:375-397)
  for any amplitudes and grid (the chain is linear and real-valued on real input);
* the Taylor-degree rule returns the smallest degree whose remainder bound meets the tolerance, monotonically in the norm."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import two_qubit as otq, vector_model
from oracle.state_transfer import halving_flags, ordered_product


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 5000), world=st.integers(1, 16))
def test_shard_bounds_is_a_balanced_contiguous_partition(n, world):
    from quantum_optimal_control._engine import shard_bounds
    blocks = [shard_bounds(n, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
    sizes = [hi - lo for lo, hi in blocks]
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=60, deadline=None)
@given(m=st.integers(1, 45), seed=st.integers(0, 2 ** 31 - 1))
def test_halving_product_tree_equals_the_time_ordered_product(m, seed):
    rng = np.random.default_rng(seed)
    ops = rng.normal(size=(m, 3, 3)) + 1j * rng.normal(size=(m, 3, 3))
    ops /= np.linalg.norm(ops, axis=(1, 2), keepdims=True)             # keep the product O(1)
    seq = np.eye(3, dtype=complex)
    for k in range(m):
        seq = ops[k] @ seq                                             # later slice on the left
    tree = ordered_product(ops, halving_flags(m))
    assert np.abs(tree - seq).max() < 1e-12 * max(1.0, np.abs(seq).max()) + 1e-14


@settings(max_examples=25, deadline=None)
@given(nt=st.integers(8, 90), pad=st.integers(0, 9), seed=st.integers(0, 2 ** 31 - 1))
def test_filter_taps_equal_the_reference_fft_chain_on_random_grids(nt, pad, seed):
    from quantum_optimal_control.two_qubit.propagator_vl import filter_impulse_response
    dt = 648e-9 / 530
    prop = otq.PropagatorVL(4, nt, pad, 50e6, dt, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6,
                            2 * np.pi * 100e6, otq.GAMMAS_CZ)
    reg = np.random.default_rng(seed).normal(size=(nt, 5))
    want = prop._filter_amplitudes(reg)                                # TQ:375-397 restated
    _, taps = filter_impulse_response(nt, pad, 50e6, dt)
    m = nt + 2 * pad
    got = np.zeros((m, 5))
    for k in range(m):
        for t in range(nt):
            got[k] += taps[(k - t - pad) % m] * reg[t]
    assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())


@settings(max_examples=200, deadline=None)
@given(nrm=st.floats(1e-4, 3.9), tol_exp=st.integers(-60, -20))
def test_taylor_degree_is_the_smallest_degree_meeting_the_remainder_bound(nrm, tol_exp):
    tol = 2.0 ** tol_exp
    q = vector_model.taylor_degree(nrm, tol=tol)
    rem = lambda d: nrm ** (d + 1) / math.factorial(d + 1)             # noqa: E731
    if q < 40:
        assert rem(q) <= tol * (1 + 1e-12)
        assert q == 2 or rem(q - 1) > tol * (1 - 1e-12)
    assert vector_model.taylor_degree(nrm * 1.5, tol=tol) >= q
