"""GPU parity at the STATED sizes of the BASELINE configs that have no reference source (configs 4-5: blockade chains) and
of the two things round 1 left unproven on the real kernels: sample sharding (`qocvl_set_shard`) and the reference's
published optimisation result (notebook 01 cell 9 / README.md:34).

Fixtures: tests/golden/cfg4_chain144_full.npz, cfg5_chain377_members.npz -- generated once by
tests/golden/make_golden_large.py from the DENSE oracle (2n x 2n expm per slice + torch autograd); they are oracle outputs
(these models do not exist in the reference), and the test names say so.
Tolerances: north_star -- fidelity 1e-10 absolute, gradients 1e-8 relative (max-norm per array)."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore", category=UserWarning)

from oracle import ensemble, state_transfer as ost, two_qubit as otq  # noqa: E402


def rel(x, y):
    x, y = np.asarray(x), np.asarray(y)
    return np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)


def np_(t):
    return t.detach().cpu().numpy()


def make_chain(nt, dt, n_atoms):
    from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL
    return PropagatorVL(16, nt, dt, n_atoms, 2 * np.pi * 2e6, 2 * np.pi * 4e6)


def test_cfg4_full_size_matches_dense_oracle_fixture(golden_dir):
    """BASELINE configs[3] as stated: 10-atom chain, n = 144, M = 4000 slices, single sample (time-parallel path of the
    samples-minor kernel) against the dense-oracle fixture: cost, metrics and the 3N-gradient."""
    gold = np.load(os.path.join(golden_dir, "cfg4_chain144_full.npz"))
    prop = make_chain(4000, 1e-9, 10)
    assert prop.dim == 144
    a, b, c = gold["a"], gold["b"], gold["c"]
    out = np_(prop.cost_and_grad_packed(prop._pack(a, b, c)))
    prop._plan.check()
    assert prop._plan.chain_info()["path"] == "samples_minor_time_parallel"
    assert abs(out[1] - float(gold["infid"])) < 1e-10
    assert abs(out[0] - float(gold["cost"])) < 1e-10 and abs(out[2] - float(gold["adiab"])) < 1e-10
    g = out[3:].reshape(3, 16, 2)
    for mine, key in zip(g, ("ga", "gb", "gc")):
        assert rel(mine, gold[key]) < 1e-8, key
    # the sequential sweep of the same kernel (no chunking) gives the same numbers
    import qocvl
    with qocvl.options(tp=0):
        seq = make_chain(4000, 1e-9, 10)
    out_seq = np_(seq.cost_and_grad_packed(seq._pack(a, b, c)))
    assert np.abs(out_seq[:3] - out[:3]).max() < 1e-12 and rel(out_seq[3:], out[3:]) < 1e-9


def test_cfg5_full_size_members_match_dense_oracle_fixture_including_gradient(golden_dir):
    """BASELINE configs[4] per-seed problem as stated: 12-atom chain, n = 377, M = 256, members of the 512-sample noise
    ensemble.  The ensemble kernel (8 samples per CTA) on the two members the fixture holds: mean cost, metrics AND
    3N-gradient against the dense oracle (round 1 only checked the cost at n = 377, at M = 12)."""
    gold = np.load(os.path.join(golden_dir, "cfg5_chain377_members.npz"))
    a, b, c = gold["a"], gold["b"], gold["c"]
    prop = make_chain(256, 2e-6 / 256, 12)
    assert prop.dim == 377
    n = len(gold["cost"])
    prop.set_ensemble(del_total=gold["del_total"], rabi_scale=gold["rabi_scale"])
    out = np_(prop.cost_and_grad_packed(prop._pack(a, b, c)))
    prop._plan.check()
    assert prop._plan.chain_info()["rows"] == 377
    assert abs(out[1] - gold["infid"].mean()) < 1e-10
    assert abs(out[0] - gold["cost"].mean()) < 1e-10 and abs(out[2] - gold["adiab"].mean()) < 1e-10
    g = out[3:].reshape(3, 16, 2)
    for mine, key in zip(g, ("ga", "gb", "gc")):
        assert rel(mine, gold[key].mean(axis=0)) < 1e-8, key
    # member by member (single-sample path of the same kernel family)
    for s in range(n):
        prop.set_ensemble(del_total=gold["del_total"][s:s + 1], rabi_scale=gold["rabi_scale"][s:s + 1])
        one = np_(prop.cost_and_grad_packed(prop._pack(a, b, c)))
        assert abs(one[0] - gold["cost"][s]) < 1e-10 and abs(one[1] - gold["infid"][s]) < 1e-10
        assert rel(one[3:].reshape(3, 16, 2)[0], gold["ga"][s]) < 1e-8


@pytest.mark.parametrize("opts", [{}, {"duo": 0}], ids=["two_lane_kernels", "lane_group_kernel"])
def test_sample_sharding_on_the_real_plan_two_halves_equal_the_whole(opts):
    """qocvl_set_shard / global_samples on the REAL kernels: a 13-sample ensemble evaluated as two 'virtual ranks'
    (7 + 6 samples, each told the global count 13) -- the sum of the two partial results is what the single all-reduce
    of a 2-GPU run produces -- must equal the unsharded evaluation.  (The torch.distributed plumbing around this is
    covered by the 2-rank gloo test; this is the part it stubs out.)"""
    import qocvl
    nominal, a, b, c = otq.notebook02_setup(seed=5, n_gauss=8, nt=120)
    S = 13
    dl, rs = ensemble.robust_cz_noise(S)
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    args = (8, 120, nominal.padding, 50e6, nominal.delta_t, 0.0, nominal.Vint, nominal.Delta_i_val, nominal.Rabi_i_val,
            nominal.Rabi_r_val, otq.GAMMAS_CZ)
    with qocvl.options(**opts):
        whole, part = PropagatorVL(*args), PropagatorVL(*args)
    whole.set_ensemble(del_total=dl, rabi_scale=rs)
    full = np_(whole.cost_and_grad_packed(whole._pack(a, b, c))).copy()
    from quantum_optimal_control._engine import shard_bounds
    acc = np.zeros_like(full)
    for rank in range(2):
        lo, hi = shard_bounds(S, 2, rank)
        mul, add = part._noise_tables(np.asarray(dl[lo:hi], float), np.asarray(rs[lo:hi], float))
        part._plan.set_noise(mul, add, global_samples=S)           # exactly what set_ensemble(shard=True) does per rank
        acc += np_(part.cost_and_grad_packed(part._pack(a, b, c)))
        part._plan.check()
    assert np.abs(acc[:3] - full[:3]).max() < 1e-14
    assert rel(acc[3:], full[3:]) < 1e-12


def test_notebook01_optimisation_reaches_the_published_fidelity():
    """The reference's own published, gradient-dependent result: notebook 01 cell 9 runs 500 plain gradient-descent steps
    at learn_rate 0.1 from the np.random.seed(1) initialisation, and README.md:34 reports a state-transfer fidelity of
    ~99.92 %.  The same loop on the engine (prop.optimize: CUDA-graph replay of cost + gradient + update) must descend
    monotonically on average and end at a fidelity consistent with that claim.  LOOSE by necessity: the notebook's stored
    prints come from a re-run that started at an already optimised point, the ARC decay constants are
    guessed, and the README gives two digits -- so the assertion is fidelity >= 99.92 % after the notebook's 500 steps
    (observed: see profiles/README.md) and a cost below the notebook's stored first print (0.00095) after 1500."""
    po, a, b, c = ost.notebook01_setup()
    from quantum_optimal_control.state_transfer.propagator_vl import PropagatorVL
    prop = PropagatorVL(po.input_dim, po.no_of_steps, po.delta_t, po.del_total, po.Delta_1, po.Rabi_1, po.Rabi_2,
                        po.Gamma_r, po.Gamma_i)
    hist, a1, b1, c1 = prop.optimize(a, b, c, num_iters=500, lr=0.1)
    hist = np_(hist)
    assert abs(hist[0] - 0.8022425853500806) < 5e-9                 # notebook 01 cell 5
    infid, adiab = (float(x) for x in prop.metrics(a1, b1, c1))
    assert hist[-1] < 0.02 * hist[0] and np.all(np.diff(hist[::50]) < 0)
    assert 1.0 - infid >= 0.9992, (infid, adiab)             # README.md:34 "fidelity ~ 99.92 %" (observed 99.989 %)
    hist2, a2, b2, c2 = prop.optimize(a1, b1, c1, num_iters=1000, lr=0.1)
    infid2, _ = (float(x) for x in prop.metrics(a2, b2, c2))
    assert float(hist2[-1]) < 0.00095 and 1.0 - infid2 >= 0.999
