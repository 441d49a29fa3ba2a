"""GPU parity tests of (i) the rows-strided big-n chain kernel and (ii) the blockade-chain (PXP) model of
BASELINE.json configs 4-5, against the dense torch oracle (oracle/blockade_chain.py) and against the
lane-per-row kernel."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore", category=UserWarning)

from oracle import blockade_chain as obc, ensemble, two_qubit as otq  # noqa: E402


def rel(x, y):
    x, y = np.asarray(x), np.asarray(y)
    return np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)


def np_(t):
    return t.detach().cpu().numpy()


def make_chain(po):
    from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL
    return PropagatorVL(po.input_dim, po.no_of_steps, po.delta_t, po.n_atoms, po.Rabi_max, po.Delta_max)


def force_big():
    """Plans created inside the block run every problem on the rows-strided kernel (qocvl_set_option "force_big")."""
    import qocvl
    return qocvl.options(force_big=1)


@pytest.mark.parametrize("n_atoms,expect_dim", [(4, 8), (5, 13), (7, 34)])
def test_chain_cost_and_gradient_match_dense_oracle(n_atoms, expect_dim):
    po, a, b, c = obc.chain_setup(n_atoms=n_atoms, n_gauss=6, nt=64)
    prop = make_chain(po)
    assert prop.dim == po.dim == expect_dim
    ref = po.value_and_grad(a, b, c)
    cost, (ga, gb, gc) = prop.target_and_grad(a, b, c)
    infid, adiab = prop.metrics(a, b, c)
    prop._plan.check()
    assert abs(float(infid) - ref[1]) < 1e-10 and abs(float(adiab) - ref[2]) < 1e-10
    assert abs(float(cost) - ref[0]) < 1e-10
    for mine, r in ((ga, ref[3]), (gb, ref[4]), (gc, ref[5])):
        assert tuple(mine.shape) == (6, 2) and rel(np_(mine), r) < 1e-8
    wave = np_(prop.return_physical_amplitudes(a, b, c))
    assert np.abs(wave - po.waves(*(torch.tensor(x) for x in (a, b, c))).numpy()).max() < 1e-13


def test_chain_dim144_matches_dense_oracle():
    """cfg 4 Hilbert space (10 atoms, blockade subspace of dimension 144) at a reduced slice count."""
    po, a, b, c = obc.chain_setup(n_atoms=10, n_gauss=16, nt=40, seed=7)
    prop = make_chain(po)
    assert prop.dim == 144
    ref = po.value_and_grad(a, b, c)
    cost, (ga, gb, gc) = prop.target_and_grad(a, b, c)
    prop._plan.check()
    assert abs(float(cost) - ref[0]) < 1e-10
    for mine, r in ((ga, ref[3]), (gb, ref[4]), (gc, ref[5])):
        assert rel(np_(mine), r) < 1e-8


def test_chain_ensemble_mean():
    po, a, b, c = obc.chain_setup(n_atoms=5, n_gauss=5, nt=48)
    rng = np.random.default_rng(3)
    dl, rs = rng.normal(0, 2 * np.pi * 0.1e6, 3), 1 + rng.normal(0, 0.01, 3)
    acc = None
    for d, s in zip(dl, rs):
        member, *_ = obc.chain_setup(n_atoms=5, n_gauss=5, nt=48, del_offset=float(d), rabi_scale=float(s))
        vals = [np.asarray(x, float) for x in member.value_and_grad(a, b, c)]
        acc = vals if acc is None else [x + y for x, y in zip(acc, vals)]
    ref = [x / 3 for x in acc]
    prop = make_chain(po)
    prop.set_ensemble(del_total=dl, rabi_scale=rs)
    out = np_(prop.cost_and_grad_packed(prop._pack(a, b, c)))
    prop._plan.check()
    assert np.abs(out[:3] - np.array(ref[:3])).max() < 1e-10
    assert rel(out[3:13], ref[3].ravel()) < 1e-8


def test_big_kernel_equals_lane_kernel_on_cz():
    """Same CZ problem through both chain kernels (the big one forced): independent data layouts, same numbers."""
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    args = (6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
    dl, rs = ensemble.robust_cz_noise(5)
    small = PropagatorVL(*args)
    small.set_ensemble(del_total=dl, rabi_scale=rs)
    out_small = np_(small.cost_and_grad_packed(small._pack(a, b, c))).copy()
    with force_big():
        big = PropagatorVL(*args)
        big.set_ensemble(del_total=dl, rabi_scale=rs)
        out_big = np_(big.cost_and_grad_packed(big._pack(a, b, c))).copy()
        big._plan.check()
    assert np.abs(out_big[:3] - out_small[:3]).max() < 1e-13
    assert rel(out_big[3:], out_small[3:]) < 1e-10


@pytest.mark.parametrize("n_atoms,nt", [(7, 16), (8, 21), (10, 6)])
def test_dense_api_for_large_n_block_triangular_dmma_path(n_atoms, nt):
    """exponentials() / propagate() (ST:247-283 / TQ:460-535) for n > 16 (n = 34, 55, 144): dense_blocks.cu keeps every
    matrix as its (u, w) block pair -- 3 n x n products instead of 8 -- and multiplies on FP64 tensor cores
    (k_zgemm_dmma).  Against the dense oracle's Pade-13 exponentials and the reference's halving product order (odd
    slice counts fold the leftover), and against the engine's own cost: chain cost == cost read off the dense propagator."""
    import torch
    from oracle.state_transfer import halving_flags, ordered_product
    po, a, b, c = obc.chain_setup(n_atoms=n_atoms, n_gauss=4, nt=nt)
    prop = make_chain(po)
    n = po.dim
    assert 2 * n > 32
    ex = np_(prop.exponentials(a, b, c))
    w = po.waves(torch.tensor(a), torch.tensor(b), torch.tensor(c))
    want = po._dense_steps(w, 0, nt).numpy()
    assert ex.shape == want.shape == (nt, 2 * n, 2 * n)
    assert np.abs(ex - want).max() < 1e-12
    assert np.abs(ex[:, n:, :n]).max() == 0.0 and np.abs(ex[:, :n, :n] - ex[:, n:, n:]).max() == 0.0
    vl = np_(prop.propagate(a, b, c))
    assert np.abs(vl - ordered_product(want, halving_flags(nt))).max() < 1e-11
    out = np_(prop.cost_and_grad_packed(prop._pack(a, b, c)))
    final = vl[:n, po.i_g]
    infid = 1.0 - abs(final[po.i_z]) ** 2
    adiab = (1.0 - np.vdot(final, vl[:n, n + po.i_g]) / po.duration).real
    assert abs(out[1] - infid) < 1e-11 and abs(out[2] - adiab) < 1e-11


def env(**kv):
    """env(QOCVL_TP=0) -> qocvl.options(tp=0): explicit plan options for the plans created inside the block (the names
    keep the historical QOCVL_ prefix in the test bodies; the library reads no environment variables)."""
    import qocvl
    return qocvl.options(**{k.replace("QOCVL_", "").lower(): v for k, v in kv.items()})


@pytest.mark.parametrize("chunk", [None, 2, 7, 64])
def test_time_parallel_path_equals_sequential_sweep(chunk):
    """Single-sample evaluations cut the slice axis into concurrent chunks (chunk propagators -> combine ->
    per-chunk sweeps); the result must equal the plain sequential sweep (QOCVL_TP=0) for any chunk length,
    including one that does not divide M (M = 81) and one larger than M/2."""
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    args = (6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
    with env(QOCVL_TP=0):
        seq = PropagatorVL(*args)
        out_seq = np_(seq.cost_and_grad_packed(seq._pack(a, b, c))).copy()
        cost_seq = float(seq.target(a, b, c))
    kv = {} if chunk is None else {"QOCVL_TP_CHUNK": chunk}
    with env(**kv):
        tp = PropagatorVL(*args)
        out_tp = np_(tp.cost_and_grad_packed(tp._pack(a, b, c))).copy()
        cost_tp = float(tp.target(a, b, c))           # cost-only path (propagators + combine)
        tp._plan.check()
    assert np.abs(out_tp[:3] - out_seq[:3]).max() < 1e-13 and abs(cost_tp - cost_seq) < 1e-13
    assert rel(out_tp[3:], out_seq[3:]) < 1e-10


def test_time_parallel_state_transfer():
    from oracle import state_transfer as ost
    from quantum_optimal_control.state_transfer.propagator_vl import PropagatorVL
    po, a, b, c = ost.notebook01_setup()
    args = (po.input_dim, po.no_of_steps, po.delta_t, po.del_total, po.Delta_1, po.Rabi_1, po.Rabi_2, po.Gamma_r, po.Gamma_i)
    with env(QOCVL_TP=0):
        seq = PropagatorVL(*args)
        out_seq = np_(seq.cost_and_grad_packed(seq._pack(a, b, c))).copy()
    tp = PropagatorVL(*args)
    out_tp = np_(tp.cost_and_grad_packed(tp._pack(a, b, c))).copy()
    assert np.abs(out_tp[:3] - out_seq[:3]).max() < 1e-13
    assert rel(out_tp[3:], out_seq[3:]) < 1e-9


@pytest.mark.parametrize("nofold", [False, True])
def test_reduced_system_equals_lane_kernel_on_cz(nofold):
    """chain_aug.cu (reachable sectors + orbit folding, 12 rows per sample; 15 without folding) against the
    full lane-per-row kernel (3 vectors x 16 rows) on a small CZ ensemble, with a ragged sample count."""
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    args = (6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
    dl, rs = ensemble.robust_cz_noise(37)
    with env(QOCVL_AUG=0):
        full = PropagatorVL(*args)
        full.set_ensemble(del_total=dl, rabi_scale=rs)
        out_full = np_(full.cost_and_grad_packed(full._pack(a, b, c))).copy()
        assert not full._plan.reduced_rows
    kv = {"QOCVL_NOFOLD": 1} if nofold else {}
    with env(**kv):
        red = PropagatorVL(*args)
        red.set_ensemble(del_total=dl, rabi_scale=rs)
        out_red = np_(red.cost_and_grad_packed(red._pack(a, b, c))).copy()
        cost_only = float(red.target(a, b, c))
        red._plan.check()
        assert red._plan.reduced_rows == (15 if nofold else 12)
    assert np.abs(out_red[:3] - out_full[:3]).max() < 1e-13 and abs(cost_only - out_full[0]) < 1e-13
    assert rel(out_red[3:], out_full[3:]) < 1e-10


def test_reduced_system_state_transfer_ensemble():
    """Same check on the 5-level state-transfer model (one start ket, its W psi partner, no symmetry hint)."""
    from oracle import state_transfer as ost
    from quantum_optimal_control.state_transfer.propagator_vl import PropagatorVL
    po, a, b, c = ost.notebook01_setup()
    args = (po.input_dim, po.no_of_steps, po.delta_t, po.del_total, po.Delta_1, po.Rabi_1, po.Rabi_2, po.Gamma_r, po.Gamma_i)
    rng = np.random.default_rng(5)
    dl, rs = rng.normal(0, 2 * np.pi * 1e5, 9), 1 + rng.normal(0, 0.01, 9)
    outs = []
    for flag in (0, 1):
        with env(QOCVL_AUG=flag):
            prop = PropagatorVL(*args)
            prop.set_ensemble(del_total=dl, rabi_scale=rs)
            outs.append(np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy())
            prop._plan.check()
            assert bool(prop._plan.reduced_rows) == bool(flag)
    assert np.abs(outs[1][:3] - outs[0][:3]).max() < 1e-13
    assert rel(outs[1][3:], outs[0][3:]) < 1e-10


@pytest.mark.parametrize("kind", ["two_classes", "offdiag_additive", "too_many_classes"])
def test_reduced_system_with_general_noise_tables(kind):
    """The reduced kernel groups generators into multiplier classes and keeps additive per-sample terms per entry.
    Noise tables the reference-style ensembles never produce: (i) a multiplier on the Delta_i generator, which gives
    the diagonal two classes; (ii) additive complex terms on the Rabi generators (off-diagonal entries);
    (iii) an independent multiplier on every generator (> 2 classes per entry: the engine must fall back to the full
    lane-per-row kernel).  Reference = the same plan API with QOCVL_AUG=0."""
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    args = (6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
    rng = np.random.default_rng(21)
    S, J = 11, 11
    mul, add = np.ones((S, J)), np.zeros((S, J), complex)
    if kind == "two_classes":
        mul[:, 2] = 1 + 0.05 * rng.standard_normal(S)
        mul[:, 3:7] = (1 + 0.02 * rng.standard_normal(S))[:, None]
        add[:, 1] = -1j * rng.normal(0, 2 * np.pi * 1e5, S)
    elif kind == "offdiag_additive":
        mul[:, 3:5] = (1 + 0.02 * rng.standard_normal(S))[:, None]
        mul[:, 5:7] = (1 + 0.02 * rng.standard_normal(S))[:, None]
        add[:, 3] = 1e-2 * (rng.standard_normal(S) + 1j * rng.standard_normal(S))
        add[:, 6] = 1e-2 * (rng.standard_normal(S) + 1j * rng.standard_normal(S))
    else:
        mul = 1 + 0.02 * rng.standard_normal((S, J))
    outs, rows = [], []
    for flag in (0, 1):
        with env(QOCVL_AUG=flag):
            prop = PropagatorVL(*args)
            prop._plan.set_noise(mul, add)
            outs.append(np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy())
            prop._plan.check()
            rows.append(prop._plan.reduced_rows)
    assert rows[0] == 0 and rows[1] == (0 if kind == "too_many_classes" else 12)
    assert np.abs(outs[1][:3] - outs[0][:3]).max() < 1e-13
    assert rel(outs[1][3:], outs[0][3:]) < 1e-10


@pytest.mark.parametrize("n_samples", [70, 3, 1])
def test_reduced_system_kernels_and_modes_agree(n_samples):
    """The two ensemble kernel families of the reduced stacked system -- chain_duo.cu (two lanes per coupled sector,
    registers only, checkpoints + recomputation, bulk-TMA staged slice batches; 4 warps or 1 warp per CTA) and
    chain_aug.cu (one sample per lane group; Taylor terms streamed or recomputed) -- on a ragged ensemble, plus the
    time-parallel form of the two-lane kernels: nine code paths, same numbers."""
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    po, a, b, c = otq.notebook02_setup(seed=4, n_gauss=6, nt=77)
    args = (6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
    dl, rs = ensemble.robust_cz_noise(n_samples)
    variants = [({"tp": 0}, "reduced_duo"), ({"tp": 0, "duo_wpc": 1}, "reduced_duo"),
                ({"tp": 0, "duo": 0}, "reduced_stacked_streamed"),
                ({"tp": 0, "duo": 0, "store": 0}, "reduced_stacked"),
                # time-parallel evaluation (what small shards of a strong-scaling run use): slice axis cut into chunks
                # (chunk propagators -> boundary states / cotangents -> per-chunk sweeps); 7 does not divide M = 81
                ({"tp": 0, "strong_chunks": 4}, "reduced_duo"), ({"tp": 0, "strong_chunks": 7}, "reduced_duo"),
                # ... with every basis column swept (no mirroring of the Van Loan pair's own-block propagator)
                ({"tp": 0, "strong_chunks": 4, "duo_nomirror": 1}, "reduced_duo"),
                # ... three basis columns per item in the chunk-propagator sweep (shared entry registers)
                ({"tp": 0, "strong_chunks": 4, "duo_cols": 1}, "reduced_duo"),
                # Xbar accumulated on every entry (default: not on the CZ model's diagonal, whose coefficient only depends
                # on the detuning channel that the pulse map forces to a constant, TQ:364)
                ({"tp": 0, "duo_nolean": 1}, "reduced_duo"),
                # one Taylor degree per slice for every role (default: every role -- an independent diagonal block of the
                # propagated system -- runs the degree ITS OWN norm bound asks for)
                ({"tp": 0, "duo_roleq": 0}, "reduced_duo"), ({"tp": 0, "duo_roleq": 0, "strong_chunks": 4}, "reduced_duo"),
                ({"tp": 0, "duo_roleq": 1}, "reduced_duo"),     # ... from the role's 1-norm bound only
                ({"tp": 0, "duo_roleq": 2}, "reduced_duo"),     # ... min with its 2-norm bound, no phase shift (default: shifted)
                ({"tp": 0, "duo_roleq": 2, "strong_chunks": 7}, "reduced_duo")]
    outs = []
    for kv, path in variants:
        with env(**kv):
            prop = PropagatorVL(*args)
            prop.set_ensemble(del_total=dl, rabi_scale=rs)
            outs.append(np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy())
            cost_only = float(prop.target(a, b, c))
            prop._plan.check()
            assert prop._plan.chain_info()["path"] == path, kv
            assert prop._plan.slice_chunks == kv.get("strong_chunks", 1 if n_samples > 1 or "duo" in kv else 5)   # one sample: M / 16
            assert abs(cost_only - outs[-1][0]) < 1e-14
            # per-role degrees: the {U psi, W psi} pair and U|11> get their own; never above the whole-system degree
            deg = prop._plan.role_degree_means()
            if path == "reduced_duo" and not kv.get("duo_roleq", 3) == 0:
                assert len(deg["per_role"]) == 2 and min(deg["per_role"]) < max(deg["per_role"]) <= prop._plan.taylor_cap
            else:
                assert deg["per_role"] == [] and deg["global"] >= 2
    for out, (kv, _) in zip(outs[1:], variants[1:]):
        assert np.abs(out[:3] - outs[0][:3]).max() < 1e-13, kv
        assert rel(out[3:], outs[0][3:]) < 1e-10, kv
    assert np.array_equal(outs[0], outs[1])          # CTA geometry does not change the arithmetic: bit for bit


def test_per_role_taylor_degrees_are_rigorous_and_tighter():
    """Every role (diagonal block of the propagated system) runs the Taylor degree its own norm bound asks for:
    whole-system 1-norm ("duo_roleq" = 0) >= role's 1-norm (1) >= min with the role's 2-norm bound (2: Collatz-Wielandt
    on the entrywise absolute values, k_slice_degrees) >= the same for the generator shifted by the role's phase
    (3, the default: X - i phi I, phi = midpoint of the range of Im diag over the role's states), slice by slice; and the
    bounds really bound the 2-norm of every member's (shifted) role block, computed here with NumPy from the oracle's
    generators, so the truncation rule  |X|^(q+1)/(q+1)! <= 2^-55  holds."""
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    po, a, b, c = otq.notebook02_setup(seed=4, n_gauss=6, nt=77)
    args = (6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
    dl, rs = ensemble.robust_cz_noise(9)
    degs, outs = {}, {}
    for mode in (0, 1, 2, 3):
        with env(tp=0, duo_roleq=mode):
            prop = PropagatorVL(*args)
            prop.set_ensemble(del_total=dl, rabi_scale=rs)
            outs[mode] = np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy()
            prop._plan.check()
            degs[mode] = prop._plan.role_degrees()
    assert degs[0].shape[0] == 1 and degs[1].shape == degs[2].shape == degs[3].shape == (2, degs[0].shape[1])
    assert (degs[3] <= degs[2]).all() and (degs[2] <= degs[1]).all() and (degs[1] <= degs[0]).all()
    assert degs[3].sum() < degs[2].sum() < degs[1].sum() < 2 * degs[0].sum()
    for mode in (1, 2, 3):
        assert np.abs(outs[mode][:3] - outs[0][:3]).max() < 1e-13 and rel(outs[mode][3:], outs[0][3:]) < 1e-10
    # true 2-norms of the role blocks of every ensemble member, unshifted and shifted
    from oracle import vector_model as vm

    def need(nrm, tol=2.0 ** -55):
        term, q = 1.0, 0
        while q < 40:
            q += 1
            term *= nrm / q
            if q >= 2 and term * nrm / (q + 1) <= tol:
                return q
        return 40
    r10, r11 = [4, 8, 12], [5, 6, 7, 9, 10, 11, 13, 14, 15]

    def blocks(member):
        wave, _ = vm.pulse_map(member, a, b, c, "tq")
        coef, _ = vm.coefficients_tq(member, wave)
        gu, gw = vm.split_generators(np.asarray(member.VL_Generators_ansatz))
        for k in range(coef.shape[0]):
            A = member.delta_t * np.tensordot(coef[k], gu, axes=[[0], [0]])
            B = member.delta_t * np.tensordot(coef[k], gw, axes=[[0], [0]])
            a0, b0 = A[np.ix_(r10, r10)], B[np.ix_(r10, r10)]
            yield k, np.block([[a0, np.zeros_like(a0)], [b0, a0]]), A[np.ix_(r11, r11)]
    phis = np.zeros((2, degs[0].shape[1]))                       # the engine's shift rule, on the noise-free generator
    for k, x0, x1 in blocks(po):
        for r, x in enumerate((x0, x1)):
            dg = np.diag(x).imag
            phis[r, k] = 0.5 * (dg.max() + dg.min())
    worst2, worst3, worst3_abs = np.zeros_like(degs[2]), np.zeros_like(degs[2]), np.zeros_like(degs[2])
    for d_, s_ in zip(dl, rs):
        member, *_ = otq.notebook02_setup(seed=4, n_gauss=6, nt=77, del_total=float(d_), rabi_scale=float(s_))
        for k, x0, x1 in blocks(member):
            for r, x in enumerate((x0, x1)):
                worst2[r, k] = max(worst2[r, k], need(np.linalg.norm(x, 2)))
                y = x - 1j * phis[r, k] * np.eye(x.shape[0])
                worst3[r, k] = max(worst3[r, k], need(np.linalg.norm(y, 2)))
                worst3_abs[r, k] = max(worst3_abs[r, k], need(np.linalg.norm(np.abs(y), 2)))
    # (the engine numbers its roles by size: find which is which from the mean degrees)
    order = [0, 1] if degs[2][0].mean() < degs[2][1].mean() else [1, 0]
    for r in (0, 1):
        assert (degs[2][order[r]] >= worst2[r]).all() and (degs[3][order[r]] >= worst3[r]).all()
        # ... and tight: within one degree of what the true 2-norm needs (shifted: of what the 2-norm of the entrywise
        # absolute values needs -- the bound's first step; the shifted diagonal has both signs, which | . | forgets)
        assert (degs[2][order[r]] - worst2[r]).max() <= 1 and (degs[3][order[r]] - worst3_abs[r]).max() <= 1
        assert (degs[3][order[r]] - worst3[r]).max() <= 2


@pytest.mark.parametrize("n_atoms", [2, 3, 4])
def test_two_lane_kernels_on_other_small_systems(n_atoms):
    """The cut search of chain_duo.cu on systems other than the reference's two models: blockade chains of 2-4 atoms
    (n = 3, 5, 8).  n = 3 is a star graph (its {U psi, W psi} pair needs a row order the search has to find); n = 5 and
    n = 8 have sectors of more than 6 rows and must fall back.  Whatever path is chosen agrees with the lane-group kernel
    and with the dense oracle."""
    po, a, b, c = obc.chain_setup(n_atoms=n_atoms, n_gauss=4, nt=40)
    rng = np.random.default_rng(5)
    dl, rs = rng.normal(0, 2 * np.pi * 0.1e6, 7), 1 + rng.normal(0, 0.01, 7)
    outs, paths = [], []
    for kv in ({}, {"duo": 0}, {"strong_chunks": 3}):
        with env(**kv):
            prop = make_chain(po)
            prop.set_ensemble(del_total=dl, rabi_scale=rs)
            outs.append(np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy())
            prop._plan.check()
            paths.append(prop._plan.chain_info()["path"])
    assert paths[0] == ("reduced_duo" if n_atoms == 2 else paths[1])
    for out in outs[1:]:
        assert np.abs(out[:3] - outs[0][:3]).max() < 1e-13 and rel(out[3:], outs[0][3:]) < 1e-10
    acc = None
    for d, s in zip(dl, rs):
        member, *_ = obc.chain_setup(n_atoms=n_atoms, n_gauss=4, nt=40, del_offset=float(d), rabi_scale=float(s))
        vals = [np.asarray(x, float) for x in member.value_and_grad(a, b, c)]
        acc = vals if acc is None else [x + y for x, y in zip(acc, vals)]
    ref = [x / len(dl) for x in acc]
    assert np.abs(outs[0][:3] - np.array(ref[:3])).max() < 1e-10
    assert rel(outs[0][3:11], ref[3].ravel()) < 1e-8


def test_profiling_hooks_of_the_two_lane_kernels():
    """qocvl_last_stage_ms / qocvl_flops_per_call modes / qocvl_hbm_bytes_per_call / qocvl_slice_chunks on a small CZ
    ensemble: the per-kernel event timings are positive and nest inside the chain's, the flop models are ordered
    (executed >= M2 on every entry >= needed >= the reverse sweep's share > 0), and the design bytes are the checkpoints
    (written once, read once) plus the per-warp gradient partials."""
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    po, a, b, c = otq.notebook02_setup(seed=4, n_gauss=6, nt=77)
    prop = PropagatorVL(6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val,
                        otq.GAMMAS_CZ)
    dl, rs = ensemble.robust_cz_noise(40)
    prop.set_ensemble(del_total=dl, rabi_scale=rs)
    plan = prop._plan
    plan.set_timing(True)
    for _ in range(2):
        prop.cost_and_grad_packed(prop._pack(a, b, c))
    plan.check()
    assert plan.chain_info()["path"] == "reduced_duo" and plan.slice_chunks == 1
    total, fwd, rev = plan.last_chain_ms(), plan.last_stage_ms(0), plan.last_stage_ms(1)
    assert fwd > 0 and rev > 0 and fwd + rev <= total * 1.001
    executed, needed, rev_share, m2 = (plan.flops_per_call(m) for m in (1, 2, 3, 4))
    assert executed > m2 > needed > rev_share > 0
    M, warps = 81, 3                                     # 40 samples = 3 warps of 16 per role
    ckpt = 2 * warps * M * 96 * 16                       # 2 roles x warps x slices x (3 rows x 32 lanes) complex
    assert plan.hbm_bytes_per_call(True) >= 2 * ckpt and plan.hbm_bytes_per_call(True) < 2 * ckpt * 1.5


def _chain_ensemble(n_atoms, nt, n_samples, seed=11, n_gauss=6):
    po, a, b, c = obc.chain_setup(n_atoms=n_atoms, n_gauss=n_gauss, nt=nt, seed=seed)
    rng = np.random.default_rng(seed + 1)
    dl, rs = rng.normal(0, 2 * np.pi * 0.1e6, n_samples), 1 + rng.normal(0, 0.01, n_samples)
    return po, a, b, c, dl, rs


def _run_chain(po, a, b, c, dl, rs, **kv):
    with env(**kv):
        prop = make_chain(po)
        if dl is not None:
            prop.set_ensemble(del_total=dl, rabi_scale=rs)
        out = np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy()
        cost_only = float(prop.target(a, b, c))
        prop._plan.check()
        info = prop._plan.chain_info()
    return out, cost_only, info


@pytest.mark.parametrize("n_atoms,n_samples", [(6, 11), (7, 19), (7, 3), (10, 9), (10, 1)])
def test_samples_minor_kernel_equals_one_cta_per_sample_kernel(n_atoms, n_samples):
    """chain_wide.cu (8 samples per CTA, samples-minor vectors, generator-by-generator products, streamed Taylor
    terms) against chain_big.cu (one CTA per sample, assembled CSR values, per-entry Xbar) on blockade-chain
    ensembles with a ragged last CTA; a single sample runs the one-sample-per-CTA variant of the same kernel."""
    po, a, b, c, dl, rs = _chain_ensemble(n_atoms, 40, n_samples)
    if n_samples == 1:
        dl = rs = None
    out_old, cost_old, info_old = _run_chain(po, a, b, c, dl, rs, QOCVL_WIDE=0)
    out_new, cost_new, info_new = _run_chain(po, a, b, c, dl, rs)
    assert info_old["path"] == "rows_strided" and info_new["path"] == "samples_minor_streamed"
    assert info_new["lanes"] == (8 if n_samples >= 5 else 1)
    assert np.abs(out_new[:3] - out_old[:3]).max() < 1e-13 and abs(cost_new - out_old[0]) < 1e-13
    assert rel(out_new[3:], out_old[3:]) < 1e-10


@pytest.mark.parametrize("threads", [256, 384, 512])
def test_samples_minor_kernel_streamed_equals_recomputed_and_any_geometry(threads):
    """STORE mode against checkpoint + recompute mode, for every CTA size the kernel is compiled for (different rows
    per thread and warps per CTA): identical sums up to the order of the per-warp reductions."""
    po, a, b, c, dl, rs = _chain_ensemble(8, 33, 13, seed=5)          # n = 55
    ref, cost_ref, _ = _run_chain(po, a, b, c, dl, rs, QOCVL_WIDE=0)
    for store in (1, 0):
        out, cost, info = _run_chain(po, a, b, c, dl, rs, QOCVL_WIDE_STORE=store, QOCVL_WIDE_THREADS=threads)
        assert info["path"] == ("samples_minor_streamed" if store else "samples_minor_recomputed")
        assert np.abs(out[:3] - ref[:3]).max() < 1e-13 and abs(cost - ref[0]) < 1e-13
        assert rel(out[3:], ref[3:]) < 1e-10


@pytest.mark.parametrize("n_atoms,nt,n_samples", [(7, 40, 11), (10, 40, 9), (8, 100, 1), (12, 12, 8)])
def test_samples_minor_kernel_phase_shift_and_two_norm_degrees(n_atoms, nt, n_samples):
    """k_wide_degrees: the sweeps of the samples-minor kernel run the Taylor series of X_k - i phi_k I (phi_k = midpoint
    of the range of Im diag X_k; the accumulated phase goes into the rotated target kets) at the degree the 2-norm bound
    of the SHIFTED generator asks for.  Same numbers as the serial 1-norm rule without shift ("wide_norms" = 0) and as the
    unshifted 2-norm rule ("wide_shift" = 0), fewer Taylor terms; ensembles, one sample (time-parallel form), n = 377."""
    po, a, b, c, dl, rs = _chain_ensemble(n_atoms, nt, n_samples, seed=3, n_gauss=4 if n_atoms == 12 else 6)
    if n_samples == 1:
        dl = rs = None
    outs, degs = [], []
    for kv in ({"QOCVL_WIDE_NORMS": 0}, {"QOCVL_WIDE_SHIFT": 0}, {}):
        with env(**kv):
            prop = make_chain(po)
            if dl is not None:
                prop.set_ensemble(del_total=dl, rabi_scale=rs)
            for _ in range(2):                                  # (the second evaluation runs with k_coefficients' norm loop off)
                out = np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy()
            cost_only = float(prop.target(a, b, c))
            prop._plan.check()
            assert prop._plan.chain_info()["path"].startswith("samples_minor")
            assert abs(cost_only - out[0]) < 1e-13
            outs.append(out)
            degs.append(prop._plan.role_degrees()[0])
    for out in outs[1:]:
        assert np.abs(out[:3] - outs[0][:3]).max() < 1e-12 and rel(out[3:], outs[0][3:]) < 1e-9
    assert (degs[2] <= degs[1]).all() and (degs[1] <= degs[0]).all() and degs[2].sum() < degs[0].sum()
    # rigour: the degree covers the true 2-norm of every member's shifted generator
    def need(nrm, tol=2.0 ** -55):
        term, q = 1.0, 0
        while q < 40:
            q += 1
            term *= nrm / q
            if q >= 2 and term * nrm / (q + 1) <= tol:
                return q
        return 40
    if n_atoms <= 8:
        w = po.waves(*(torch.tensor(x) for x in (a, b, c))).numpy()
        hx, nn = po.h_drive.real, po.h_number.real
        n, M = po.dim, po.no_of_steps
        for d_, s_ in zip(dl if dl is not None else [0.0], rs if rs is not None else [1.0]):
            for k in range(M):
                ham0 = po.Rabi_max * w[k, 0] * hx - po.Delta_max * w[k, 1] * nn          # nominal: defines the shift
                dg = -np.diag(ham0) * po.delta_t                                        # Im diag of X = -i dt H
                phi = 0.5 * (dg.max() + dg.min())
                ham = s_ * po.Rabi_max * w[k, 0] * hx - (po.Delta_max * w[k, 1] + d_) * nn
                pen = np.zeros((n, n)); pen[po.i_g, po.i_g] = 1 - k / M; pen[po.i_z, po.i_z] = k / M
                y = -1j * po.delta_t * ham - 1j * phi * np.eye(n)
                x = np.block([[y, po.delta_t * pen], [np.zeros((n, n)), y]])
                assert degs[2][k] >= need(np.linalg.norm(x, 2)), k


def test_samples_minor_kernel_dim377_against_dense_oracle_sample():
    """cfg 5 Hilbert space (12 atoms, n = 377) on a short grid: ensemble of 8 = one full CTA; the ensemble mean must
    equal the mean of the dense torch oracle evaluated member by member for the cost, and the one-CTA-per-sample
    kernel for the gradient."""
    po, a, b, c, dl, rs = _chain_ensemble(12, 12, 8, seed=2, n_gauss=4)
    out_new, _, info = _run_chain(po, a, b, c, dl, rs)
    assert info["rows"] == 377 and info["path"] == "samples_minor_streamed"
    out_old, _, _ = _run_chain(po, a, b, c, dl, rs, QOCVL_WIDE=0)
    acc = 0.0
    for d, s in zip(dl, rs):
        member, *_ = obc.chain_setup(n_atoms=12, n_gauss=4, nt=12, seed=2, del_offset=float(d), rabi_scale=float(s))
        acc += float(member.cost_torch(*(torch.tensor(x) for x in (a, b, c)))[0])
    assert abs(out_new[0] - acc / 8) < 1e-10
    assert np.abs(out_new[:3] - out_old[:3]).max() < 1e-13 and rel(out_new[3:], out_old[3:]) < 1e-10


def test_samples_minor_kernel_general_blocks(monkeypatch):
    """The general instantiation of chain_wide.cu: off-diagonal entries with DIFFERENT values (site-dependent Rabi
    frequencies, complex phases) and an off-diagonal pair in a top-right (penalty) block, which the blockade-chain
    model never produces.  Reference: the one-CTA-per-sample kernel on the same generators."""
    import quantum_optimal_control.blockade_chain.propagator_vl as mod
    plain = mod.chain_generators

    def perturbed(n_atoms, Rabi_max, Delta_max):
        states, gu, gw, i_g, i_z = plain(n_atoms, Rabi_max, Delta_max)
        n = len(states)
        rng = np.random.default_rng(0)
        phase = np.exp(1j * rng.uniform(-0.3, 0.3, (n, n)))
        scale = 1 + 0.2 * rng.uniform(-1, 1, (n, n))
        factor = np.triu(scale * phase, 1)
        gu[0] = gu[0] * (factor + factor.conj().T)            # Hermitian, entries all different
        gw[2][1, 3] = 0.3 + 0.1j; gw[2][3, 1] = 0.3 - 0.1j    # off-diagonal pair in a penalty block
        return states, gu, gw, i_g, i_z

    monkeypatch.setattr(mod, "chain_generators", perturbed)
    po, a, b, c, dl, rs = _chain_ensemble(7, 36, 10, seed=9)
    out_old, cost_old, info_old = _run_chain(po, a, b, c, dl, rs, QOCVL_WIDE=0)
    out_new, cost_new, info_new = _run_chain(po, a, b, c, dl, rs)
    assert info_old["path"] == "rows_strided" and info_new["path"] == "samples_minor_streamed"
    assert info_new["slots"] == 2 and info_new["classes"] == 3     # generator blocks: 2 off-diagonal, 3 diagonal
    assert np.abs(out_new[:3] - out_old[:3]).max() < 1e-13 and abs(cost_new - out_old[0]) < 1e-13
    assert rel(out_new[3:], out_old[3:]) < 1e-10
    out_rec, _, info_rec = _run_chain(po, a, b, c, dl, rs, QOCVL_WIDE_STORE=0)
    assert info_rec["path"] == "samples_minor_recomputed"
    assert np.abs(out_rec[:3] - out_old[:3]).max() < 1e-13 and rel(out_rec[3:], out_old[3:]) < 1e-10


def test_multi_start_concurrent_seeds_equal_sequential_evaluation():
    """MultiStart (several replicas of the propagator evaluating different pulse seeds concurrently on separate
    streams) must return, per seed, exactly what one propagator returns evaluating the seeds one after the other;
    its device-side descent must lower every seed's cost."""
    from quantum_optimal_control.multi_start import MultiStart
    po, a, b, c, dl, rs = _chain_ensemble(7, 30, 9, seed=4)

    def make():
        prop = make_chain(po)
        prop.set_ensemble(del_total=dl, rabi_scale=rs)
        return prop

    single = make()
    seeds = []
    for k in range(5):
        rng = np.random.default_rng(50 + k)
        seeds.append(single._pack(rng.uniform(-1, 1, (6, 2)), rng.uniform(-1, 1, (6, 2)), rng.uniform(0, 0.5, (6, 2))))
    seeds = torch.stack(seeds)
    ref = torch.stack([single.cost_and_grad_packed(x).clone() for x in seeds])
    ms = MultiStart(make, n_concurrent=3)
    out = ms.cost_and_grad(seeds)
    torch.cuda.synchronize()
    ms.check()
    assert ms.n_concurrent == 3 and len(ms.replicas) == 3
    assert torch.equal(out, ref)
    hist, final = ms.optimize(seeds, num_iters=6, lr=0.05)
    assert bool((hist[-1] < hist[0]).all())
    i, best = ms.best(hist[-1], 5)
    assert best == float(hist[-1].min()) and i == int(hist[-1].argmin())
    auto = MultiStart(make, sm_count=4)       # 2 CTAs per evaluation on a pretend 4-SM device -> 2 replicas
    auto.cost_and_grad(seeds[:2])
    assert auto.n_concurrent == 2


@pytest.mark.parametrize("chunk", [None, 7, 33, 90])
def test_samples_minor_time_parallel_single_sample(chunk):
    """A single sample on a large Hilbert space: column sweeps of every chunk (8 basis columns per CTA) -> combine ->
    per-chunk forward + reverse sweeps must equal the plain sequential sweep of the same kernel (QOCVL_TP=0) and the
    one-CTA-per-sample kernel, for chunk lengths that do not divide M = 100 and for a single short tail chunk."""
    po, a, b, c, _, _ = _chain_ensemble(8, 100, 1, seed=6)          # n = 55
    out_old, cost_old, _ = _run_chain(po, a, b, c, None, None, QOCVL_WIDE=0)
    out_seq, cost_seq, info_seq = _run_chain(po, a, b, c, None, None, QOCVL_TP=0)
    kv = {} if chunk is None else {"QOCVL_TP_CHUNK": chunk}
    out_tp, cost_tp, info_tp = _run_chain(po, a, b, c, None, None, **kv)
    assert info_seq["path"] == "samples_minor_streamed" and info_tp["path"] == "samples_minor_time_parallel"
    assert info_tp["ctas"] == -(-100 // (chunk or 8))
    for out, cost in ((out_seq, cost_seq), (out_tp, cost_tp)):
        assert np.abs(out[:3] - out_old[:3]).max() < 1e-13 and abs(cost - out_old[0]) < 1e-13
        assert rel(out[3:], out_old[3:]) < 1e-10


def test_chain_ensemble_golden_fixture(golden_dir):
    """The committed blockade-chain fixture (n = 55, 3 noise samples, dense-oracle mean): cost to 1e-10, gradients to
    1e-8 relative, through every kernel that can run it."""
    gold = np.load(os.path.join(golden_dir, "chain8_ensemble.npz"))
    po, *_ = obc.chain_setup(n_atoms=8, n_gauss=5, nt=48, seed=21)
    a, b, c = gold["a"], gold["b"], gold["c"]
    for kv in ({}, {"QOCVL_WIDE_SPC": 8}, {"QOCVL_WIDE_STORE": 0}, {"QOCVL_WIDE": 0}):
        out, cost_only, info = _run_chain(po, a, b, c, gold["del_total"], gold["rabi_scale"], **kv)
        assert abs(out[0] - float(gold["cost"])) < 1e-10 and abs(cost_only - float(gold["cost"])) < 1e-10, info
        assert abs(out[1] - float(gold["infid"])) < 1e-10 and abs(out[2] - float(gold["adiab"])) < 1e-10
        g = out[3:].reshape(3, 5, 2)
        for mine, key in ((g[0], "ga"), (g[1], "gb"), (g[2], "gc")):
            assert rel(mine, gold[key]) < 1e-8, (kv, key)


# ---- optional complex64 mode of the ensemble kernel (north_star: "batched complex128 (optional complex64)") ------------
# Stated bounds against the complex128 result (include/qocvl.h): cost, infidelity and adiabaticity
# 5e-5 absolute, every gradient array 1e-2 of its largest entry.
TOL_C64_COST = 5e-5
TOL_C64_GRAD_REL = 1e-2


def test_complex64_ensemble_within_stated_bounds_of_oracle_and_of_complex128():
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    S = 6
    dl, rs = ensemble.robust_cz_noise(S)
    members, (a, b, c) = ensemble.robust_cz_members(dl, rs, seed=5, n_gauss=8, nt=120)
    ref = ensemble.ensemble_value_and_grad(members, a, b, c)
    po, *_ = otq.notebook02_setup(seed=5, n_gauss=8, nt=120)
    args = (8, 120, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
    prop = PropagatorVL(*args)
    prop.set_ensemble(del_total=dl, rabi_scale=rs)
    full = np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy()
    outs = {}
    # two-lane kernels in fp32 (default) / lane-group kernel with streamed terms / with checkpoints + recomputation
    for flag, (kv, path) in enumerate([({}, "reduced_duo"), ({"duo": 0}, "reduced_stacked_streamed"),
                                       ({"duo": 0, "store": 0}, "reduced_stacked")]):
        with env(**kv):
            p64 = PropagatorVL(*args)
            p64.set_precision("complex64")
            assert p64._plan.precision == "complex64"
            p64.set_ensemble(del_total=dl, rabi_scale=rs)
            outs[flag] = np_(p64.cost_and_grad_packed(p64._pack(a, b, c))).copy()
            p64._plan.check()
            assert p64._plan.chain_info()["path"] == path
    n = 8 * 5
    for out in outs.values():
        assert np.abs(out[:3] - np.array(ref[:3])).max() < TOL_C64_COST          # vs the CPU oracle
        assert np.abs(out[:3] - full[:3]).max() < TOL_C64_COST                   # vs the complex128 kernel
        for i, r in enumerate(ref[3:6]):
            assert rel(out[3 + i * n:3 + (i + 1) * n], np.asarray(r).ravel()) < TOL_C64_GRAD_REL
        assert np.abs(out - full).max() > 0.0                                    # it really is another arithmetic
    # the complex64 kernels see the same terms up to fp32 rounding
    assert np.abs(outs[1][:3] - outs[0][:3]).max() < TOL_C64_COST and np.abs(outs[2][:3] - outs[1][:3]).max() < TOL_C64_COST
    # switching back gives the complex128 numbers
    p64.set_precision("complex128")
    back = np_(p64.cost_and_grad_packed(p64._pack(a, b, c))).copy()
    assert np.abs(back[:3] - full[:3]).max() < 1e-13 and rel(back[3:], full[3:]) < 1e-10


def test_complex64_single_sample_and_state_transfer():
    """One sample (the time-parallel form of the two-lane kernels, in fp32) and the 5-level model."""
    from oracle import autograd_ref, state_transfer as ost
    from quantum_optimal_control.state_transfer.propagator_vl import PropagatorVL as StVL
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL as TqVL
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    prop = TqVL(6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val,
                otq.GAMMAS_CZ)
    so, sa, sb, sc = ost.notebook01_setup()
    pst = StVL(so.input_dim, so.no_of_steps, so.delta_t, so.del_total, so.Delta_1, so.Rabi_1, so.Rabi_2, so.Gamma_r,
               so.Gamma_i)
    for engine, model, ctrl in ((prop, po, (a, b, c)), (pst, so, (sa, sb, sc))):
        engine.set_precision("complex64")
        cost, grads = engine.target_and_grad(*ctrl)
        engine._plan.check()
        assert engine._plan.chain_info()["path"] == "reduced_duo" and engine._plan.slice_chunks > 1
        ref = autograd_ref.value_and_grad(model, *ctrl)
        assert abs(float(cost) - ref[0]) < TOL_C64_COST
        for mine, r in zip(grads, ref[3:]):
            assert rel(np_(mine), r) < TOL_C64_GRAD_REL


@pytest.mark.parametrize("n_atoms,nt,n_samples,kv", [(7, 40, 19, {}), (7, 40, 19, {"QOCVL_WIDE_STORE": 0}),
                                                    (10, 40, 9, {}), (8, 100, 1, {}), (8, 100, 1, {"QOCVL_TP": 0}),
                                                    (7, 36, 3, {})])
def test_complex64_on_large_hilbert_spaces(n_atoms, nt, n_samples, kv):
    """The complex64 instantiation of the samples-minor kernel (chain_wide.cuh<cplxf>: vectors, Taylor terms, the HBM
    stream / checkpoints and the per-slice Ybar accumulators in fp32; coefficients, cost, cotangents at the chunk
    boundaries and every reduction in fp64) on blockade chains with n = 34, 55, 144: ensembles (streamed and recomputed
    terms), a ragged small ensemble on the one-sample-per-CTA variant, and ONE sample on the time-parallel and the
    sequential form -- all within the stated bounds of the complex128 result of the same kernel."""
    po, a, b, c, dl, rs = _chain_ensemble(n_atoms, nt, n_samples, seed=5)
    if n_samples == 1:
        dl = rs = None
    with env(**kv):
        prop = make_chain(po)
        if dl is not None:
            prop.set_ensemble(del_total=dl, rabi_scale=rs)
        full = np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy()
        path128 = prop._plan.chain_info()["path"]
        prop.set_precision("complex64")
        out = np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy()
        cost_only = float(prop.target(a, b, c))
        prop._plan.check()
        assert prop._plan.precision == "complex64" and prop._plan.chain_info()["path"] == path128
        assert path128.startswith("samples_minor")
        assert np.abs(out - full).max() > 0.0                               # it really is another arithmetic
        assert np.abs(out[:3] - full[:3]).max() < TOL_C64_COST and abs(cost_only - full[0]) < TOL_C64_COST
        n = 6 * 2
        for i in range(3):
            assert rel(out[3 + i * n:3 + (i + 1) * n], full[3 + i * n:3 + (i + 1) * n]) < TOL_C64_GRAD_REL
        prop.set_precision("complex128")                                    # switching back gives the complex128 numbers
        back = np_(prop.cost_and_grad_packed(prop._pack(a, b, c))).copy()
        assert np.array_equal(back, full)


def test_complex64_fails_loudly_where_no_complex64_kernel_exists():
    """The one-CTA-per-sample fallback of the large-Hilbert-space path has no complex64 instantiation: the evaluation
    raises instead of running in fp64."""
    from qocvl import QocvlError
    po, a, b, c = obc.chain_setup(n_atoms=7, n_gauss=6, nt=64)      # dim 34: q and p together exceed 32 rows
    with env(QOCVL_WIDE=0):
        prop = make_chain(po)
        prop.set_precision("complex64")
        with pytest.raises(QocvlError):
            prop.target_and_grad(a, b, c)
        with pytest.raises(ValueError):
            prop.set_precision("complex32")
        prop.set_precision("complex128")                            # and the plan is still usable afterwards
        cost, _ = prop.target_and_grad(a, b, c)
        assert abs(float(cost) - po.value_and_grad(a, b, c)[0]) < 1e-10


def test_complex64_optimiser_loop_replays_from_a_graph_and_tracks_complex128():
    """The device-side optimiser loop in complex64: graph replay equals eager launches bit for bit (the kernel is
    deterministic in either arithmetic), the cost descends, and the 12-step history stays within the stated bound of the
    complex128 history."""
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    po, a, b, c = otq.notebook02_setup(seed=8, n_gauss=6, nt=77)
    prop = PropagatorVL(6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val,
                        po.Rabi_r_val, otq.GAMMAS_CZ)
    dl, rs = ensemble.robust_cz_noise(9)
    prop.set_ensemble(del_total=dl, rabi_scale=rs)
    hist_128, *_ = prop.optimize(a, b, c, num_iters=12, lr=0.02)
    prop.set_precision("complex64")
    hist_g, ag, bg, cg = prop.optimize(a, b, c, num_iters=12, lr=0.02)
    hist_e, ae, be, ce = prop.optimize(a, b, c, num_iters=12, lr=0.02, use_graph=False)
    prop._plan.check()
    assert prop._plan.precision == "complex64" and prop._plan.chain_info()["path"] == "reduced_duo"
    assert torch.equal(hist_g, hist_e) and torch.equal(ag, ae) and torch.equal(cg, ce)
    assert float(hist_g[-1]) < float(hist_g[0])
    assert float((hist_g - hist_128).abs().max()) < TOL_C64_COST
