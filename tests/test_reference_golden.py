"""Parity against the REFERENCE'S OWN SOURCE: `tests/golden/ref_*.npz` hold outputs of
/root/reference/code/quantum_optimal_control/{state_transfer,two_qubit}/propagator_vl.py executed unmodified on the
torch-backed `jax` stand-in of tests/golden/jax_shim (generator: tests/golden/make_reference_golden.py; the gradient is
reverse-mode autodiff THROUGH the reference's code, i.e. what `jax.value_and_grad(target, argnums=(0,1,2))` computes,
TQ:592 / notebook 01 cell 9).  Nothing here reads /root/reference at run time.

* CPU tests (`-m "not gpu"`): the oracle restatement (oracle/) against these reference outputs, stage by stage --
  this is what pins the oracle to the reference beyond the single scalar of notebook 01 cell 5.
* GPU tests (`-m gpu`): the CUDA engine, through the C-ABI, against the same reference outputs, including four
  members of BASELINE configs[2] at its STATED size (N = 16, nt = 1888, pad = 56, M = 2000) through the ensemble kernel
  with streamed and with recomputed Taylor terms.

Tolerances (north_star): fidelity 1e-10 absolute, gradients 1e-8 relative (max-norm per array).  The adiabaticity term
carries ~2e-10 evaluation noise end to end (0/0-like projector weights where the filtered pulses vanish), so it and the
cost are held to 2e-9 -- the same bound tests/test_gpu_parity.py states."""
import os
import warnings

import numpy as np
import pytest

warnings.filterwarnings("ignore", category=UserWarning)

from oracle import autograd_ref, ensemble, state_transfer as ost, two_qubit as otq  # noqa: E402

TOL_FID, TOL_ADIAB, TOL_GRAD = 1e-10, 2e-9, 1e-8
NB01_STORED_COST = 0.8022425853500806          # notebook 01 cell 5, the one number the reference stores


def rel(x, y):
    x, y = np.asarray(x), np.asarray(y)
    return np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)


CASES = {
    "ref_cfg1_state_transfer": ost.notebook01_setup,
    "ref_cfg2_two_qubit": otq.notebook02_setup,
    "ref_small_two_qubit_odd": lambda: otq.notebook02_setup(seed=3, n_gauss=6, nt=77),
}


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


# ------------------------------------------------------------------------------------------------ CPU: oracle vs reference
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_the_reference_source_stage_by_stage(name, golden_dir):
    gold = load(golden_dir, name)
    po, a, b, c = CASES[name]()
    assert np.array_equal(a, gold["a"]) and np.array_equal(b, gold["b"]) and np.array_equal(c, gold["c"])
    assert np.array_equal(po.VL_Generators_ansatz, gold["generators"])                  # a6: bit-identical generators
    assert list(po.contraction_array) == list(gold["contraction_array"]) and po.duration == float(gold["duration"])
    wave = po.return_physical_amplitudes(a, b, c)
    assert np.abs(wave - gold["wave"]).max() < 1e-13                                    # a1-a4
    aux = po.return_auxiliary_amplitudes(a, b, c)
    assert np.abs(aux[:, :5] - gold["aux"][:, :5]).max() < 1e-12                        # a5 (Hamiltonian columns)
    assert np.abs(po.aux_from_wave(gold["wave"]) - gold["aux"]).max() < 1e-12           # a5 incl. projector weights
    exps = po.exponentials(a, b, c)
    assert np.abs(exps[gold["exp_index"]] - gold["exp_slices"]).max() < 5e-9            # a7-a8 (0/0 noise x dt scale)
    assert np.abs(po.propagate(a, b, c) - gold["propagator"]).max() < 5e-9              # a9
    infid, adiab = po.metrics(a, b, c)
    assert abs(infid - float(gold["infid"])) < TOL_FID and abs(adiab - float(gold["adiab"])) < TOL_ADIAB   # a10
    assert abs(po.target(a, b, c) - float(gold["cost"])) < TOL_ADIAB                                       # a11
    res = autograd_ref.value_and_grad(po, a, b, c)                                                         # a12
    for mine, key in zip(res[3:], ("ga", "gb", "gc")):
        assert rel(mine, gold[key]) < TOL_GRAD, key


def test_reference_source_reproduces_its_own_stored_number(golden_dir):
    """The shim-run reference gives notebook 01 cell 5's stored cost (ARC decay constants guessed: SURVEY 0.3 puts
    their influence at 2e-8, so 5e-9 is the assertion and 5e-11 the observed difference)."""
    gold = load(golden_dir, "ref_cfg1_state_transfer")
    assert abs(float(gold["cost"]) - NB01_STORED_COST) < 5e-9


def test_oracle_gradient_step_equals_the_reference_single_optimization_step(golden_dir):
    gold = load(golden_dir, "ref_small_two_qubit_step")                                 # TQ:580-600 run as written
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    res = autograd_ref.value_and_grad(po, a, b, c)
    lr = float(gold["lr"])
    assert abs(res[0] - float(gold["cost"])) < TOL_ADIAB
    for x, g, key in ((a, res[3], "new_a"), (b, res[4], "new_b"), (c, res[5], "new_c")):
        assert np.abs((x - lr * g) - gold[key]).max() < 1e-9


def test_oracle_matches_reference_members_of_cfg3_at_full_size(golden_dir):
    """BASELINE configs[2] per-sample problem at its stated size (M = 2000, N = 16): oracle vs reference source on
    two of the four stored members (cost, metrics, gradient, 32x32 propagator)."""
    gold = load(golden_dir, "ref_cfg3_members")
    members, (a, b, c) = ensemble.robust_cz_members(gold["del_total"][:2], gold["rabi_scale"][:2], seed=0, n_gauss=16,
                                                    nt=1888)
    assert np.array_equal(a, gold["a"]) and np.array_equal(c, gold["c"])
    for s, po in enumerate(members):
        assert po.no_of_steps + 2 * po.padding == 2000
        res = autograd_ref.value_and_grad(po, a, b, c)
        assert abs(res[1] - gold["infid"][s]) < TOL_FID and abs(res[2] - gold["adiab"][s]) < TOL_ADIAB
        assert abs(res[0] - gold["cost"][s]) < TOL_ADIAB
        for mine, key in zip(res[3:], ("ga", "gb", "gc")):
            assert rel(mine, gold[key][s]) < TOL_GRAD, key
    assert np.abs(members[0].propagate(a, b, c) - gold["propagator"][0]).max() < 5e-9


# ------------------------------------------------------------------------------------------------ GPU: engine vs reference
def make_engine(name, po):
    if name.startswith("ref_cfg1"):
        from quantum_optimal_control.state_transfer.propagator_vl import PropagatorVL
        return PropagatorVL(po.input_dim, po.no_of_steps, po.delta_t, po.del_total, po.Delta_1, po.Rabi_1, po.Rabi_2,
                            po.Gamma_r, po.Gamma_i)
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    return PropagatorVL(po.input_dim, po.no_of_steps, po.padding, 50e6, po.delta_t, po.del_total, po.Vint,
                        po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_engine_matches_the_reference_source(name, golden_dir):
    gold = load(golden_dir, name)
    po, a, b, c = CASES[name]()
    prop = make_engine(name, po)
    np_ = lambda t: t.detach().cpu().numpy()  # noqa: E731
    assert np.abs(np_(prop.return_physical_amplitudes(a, b, c)) - gold["wave"]).max() < 1e-13
    assert np.abs(np_(prop.return_auxiliary_amplitudes(a, b, c))[:, :5] - gold["aux"][:, :5]).max() < 1e-12
    assert np.abs(np_(prop.exponentials(a, b, c))[gold["exp_index"]] - gold["exp_slices"]).max() < 5e-9
    assert np.abs(np_(prop.propagate(a, b, c)) - gold["propagator"]).max() < 5e-9
    infid, adiab = (float(x) for x in prop.metrics(a, b, c))
    assert abs(infid - float(gold["infid"])) < TOL_FID and abs(adiab - float(gold["adiab"])) < TOL_ADIAB
    cost, grads = prop.target_and_grad(a, b, c)
    prop._plan.check()
    assert abs(float(cost) - float(gold["cost"])) < TOL_ADIAB
    for mine, key in zip(grads, ("ga", "gb", "gc")):
        assert rel(np_(mine), gold[key]) < TOL_GRAD, key
    if name == "ref_cfg1_state_transfer":
        assert abs(float(cost) - NB01_STORED_COST) < 5e-9


@pytest.mark.gpu
def test_engine_single_optimization_step_equals_the_reference(golden_dir):
    from quantum_optimal_control.two_qubit.propagator_vl import single_optimization_step
    gold = load(golden_dir, "ref_small_two_qubit_step")
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    prop = make_engine("tq", po)
    cost, na, nb, nc = single_optimization_step(prop, a, b, c, lr=float(gold["lr"]))
    assert abs(float(cost) - float(gold["cost"])) < TOL_ADIAB
    for mine, key in ((na, "new_a"), (nb, "new_b"), (nc, "new_c")):
        assert np.abs(mine.cpu().numpy() - gold[key]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("opts,path", [({}, "reduced_duo"), ({"strong_chunks": 1}, "reduced_duo"),
                                       ({"duo": 0}, "reduced_stacked_streamed"),
                                       ({"duo": 0, "store": 0}, "reduced_stacked")],
                         ids=["default_time_parallel", "sequential_sweeps", "lane_groups", "lane_groups_recompute"])
def test_cfg3_full_size_members_match_the_reference(opts, path, golden_dir):
    """BASELINE configs[2] at its stated per-sample size (N = 16, nt = 1888, M = 2000): the ensemble kernels on the
    four members the fixture holds -- mean cost / infidelity / adiabaticity / gradient against the mean of the reference
    source's per-member values -- on the two-lane kernels (default) and, with streamed and with recomputed Taylor terms, on the older ensemble kernels."""
    import qocvl
    gold = load(golden_dir, "ref_cfg3_members")
    nominal, a, b, c = otq.notebook02_setup(seed=0, n_gauss=16, nt=1888)
    with qocvl.options(**opts):
        prop = make_engine("tq", nominal)
    prop.set_ensemble(del_total=gold["del_total"], rabi_scale=gold["rabi_scale"])
    out = prop.cost_and_grad_packed(prop._pack(a, b, c)).cpu().numpy()
    prop._plan.check()
    info = prop._plan.chain_info()
    assert info["path"].startswith("reduced_"), info
    assert info["path"] == path, info
    assert prop._plan.slice_chunks == (16 if opts == {} else 1)   # 4 samples x 2000 slices: 16 concurrent chunks by default
    assert abs(out[1] - gold["infid"].mean()) < TOL_FID
    assert abs(out[2] - gold["adiab"].mean()) < TOL_ADIAB and abs(out[0] - gold["cost"].mean()) < TOL_ADIAB
    g = out[3:].reshape(3, 16, 5)
    for mine, key in zip(g, ("ga", "gb", "gc")):
        assert rel(mine, gold[key].mean(axis=0)) < TOL_GRAD, key
    # member by member through the single-sample (time-parallel) path
    prop.set_ensemble(del_total=gold["del_total"][:1], rabi_scale=gold["rabi_scale"][:1])
    one = prop.cost_and_grad_packed(prop._pack(a, b, c)).cpu().numpy()
    assert abs(one[1] - gold["infid"][0]) < TOL_FID and abs(one[0] - gold["cost"][0]) < TOL_ADIAB
    assert rel(one[3:].reshape(3, 16, 5)[0], gold["ga"][0]) < TOL_GRAD


@pytest.mark.gpu
def test_cfg3_dense_propagator_at_full_size_matches_the_reference(golden_dir):
    """`propagate()` (a8-a9: M = 2000 dense 32x32 exponentials + the reference's halving order) on member 0."""
    gold = load(golden_dir, "ref_cfg3_members")
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
    members, (a, b, c) = ensemble.robust_cz_members(gold["del_total"][:1], gold["rabi_scale"][:1], seed=0, n_gauss=16,
                                                    nt=1888)
    prop = make_engine("tq", members[0])
    assert isinstance(prop, PropagatorVL)
    vl = prop.propagate(a, b, c).cpu().numpy()
    assert np.abs(vl - gold["propagator"][0]).max() < 5e-9
    assert np.abs(prop.return_physical_amplitudes(a, b, c).cpu().numpy() - gold["wave"]).max() < 1e-13
