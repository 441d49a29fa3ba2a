"""CPU tests of the host logic and of the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import qocvl
from oracle import state_transfer as ost, two_qubit as otq, vector_model
from quantum_optimal_control import _engine
from quantum_optimal_control.state_transfer import propagator_vl as pst
from quantum_optimal_control.two_qubit import propagator_vl as ptq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = qocvl.load()
    header = open(os.path.join(ROOT, "include", "qocvl.h")).read()
    declared = set(re.findall(r"\b(qocvl_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed from include/qocvl.h"
    assert declared == set(qocvl.SIGNATURES), declared ^ set(qocvl.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), f"libqocvl.so does not export {name}"
    assert lib.qocvl_version() >= 100
    assert lib.qocvl_strerror(4).decode().startswith("slice generator norm")


def test_desc_struct_layout_matches_header():
    # 10 ints, 6 doubles, double[4]: the C compiler pads nothing between them (10*4 = 40, 8-aligned)
    assert ctypes.sizeof(qocvl.Desc) == 10 * 4 + 6 * 8 + 4 * 8


def test_bad_arguments_are_status_codes_not_crashes():
    lib = qocvl.load()
    assert lib.qocvl_plan_create(None, None, 0) == 1
    assert lib.qocvl_set_filter(None, None) == 1
    assert lib.qocvl_cost(None, None, None, None) == 1
    assert lib.qocvl_plan_destroy(None) == 0


def test_host_generators_match_reference_layout():
    po, *_ = ost.notebook01_setup()
    kets = pst.PropagatorVL.nLevelAtomBasis(5)
    args = [*kets, 0.3, 0.2, po.Rabi_1, po.Rabi_2, po.Gamma_r, po.Gamma_i]
    mine = pst.PropagatorVL.VL_Generators(pst.PropagatorVL.Hamiltonian(args), kets, 5)
    ref = ost.PropagatorVL.VL_Generators(ost.PropagatorVL.Hamiltonian(args), kets, 5)
    assert mine.shape == (11, 10, 10) and np.array_equal(mine, ref)
    po, *_ = otq.notebook02_setup()
    one = np.eye(4, dtype=complex)
    args = (one[0], one[1], one[2], one[3], 0.7, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val,
            *otq.GAMMAS_CZ)
    mine = ptq.PropagatorVL._VL_Generators(ptq.PropagatorVL._Hamiltonian(*args), one[0], one[1])
    ref = otq.PropagatorVL._VL_Generators(otq.PropagatorVL._Hamiltonian(*args), one[0], one[1])
    assert mine.shape == (11, 32, 32) and np.array_equal(mine, ref)


def test_filter_taps_reproduce_the_fft_filter():
    po, a, b, c = otq.notebook02_setup()
    window, taps = ptq.filter_impulse_response(po.no_of_steps, po.padding, 50e6, po.delta_t)
    assert np.array_equal(window, po.Gaussian_filter)
    reg = po._regularize_amplitudes(a, b, c)
    M, pad, nt = 530, 15, 500
    idx = (np.arange(M)[:, None] - np.arange(nt)[None, :] - pad) % M
    assert np.abs(taps[idx] @ reg - po._filter_amplitudes(reg)).max() < 1e-14


def test_contraction_array_and_shards():
    assert _engine.halving_flags(530) == ost.halving_flags(530) == [bool((530 >> s) & 1) for s in range(9)]
    assert _engine.halving_flags(1) == []
    for n, w in ((4096, 8), (10, 3), (2, 4), (7, 1)):
        spans = [_engine.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_taylor_degree_rule():
    def remainder(nrm, q):
        return nrm ** (q + 1) / float(np.prod(np.arange(1, q + 2, dtype=float)))

    for nrm in (0.003, 0.077, 0.15, 0.3, 1.15, 4.0):
        q = vector_model.taylor_degree(nrm)
        assert 2 <= q <= 40
        assert remainder(nrm, q) <= 2.0 ** -60 * 1.0001           # accurate enough ...
        assert q == 2 or remainder(nrm, q - 1) > 2.0 ** -60 * 0.9999   # ... and minimal
    assert vector_model.taylor_degree(0.3) < vector_model.taylor_degree(1.15) < vector_model.taylor_degree(4.0)


def test_no_product_code_imports_the_oracle():
    pkg = os.path.join(ROOT, "optimal-control-rydberg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_plan_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pst.PropagatorVL(10, 500, 4e-10, 0, 0, 1e9, 1e9, 1e4, 1e7)


def test_blockade_chain_generators_match_oracle():
    """Product-side generator builder of the (non-reference) blockade-chain model vs the oracle's own enumeration."""
    from oracle import blockade_chain as obc
    from quantum_optimal_control.blockade_chain.propagator_vl import blockade_basis, chain_generators
    assert [len(blockade_basis(k)) for k in (1, 2, 3, 4, 5, 10, 12)] == [2, 3, 5, 8, 13, 144, 377]   # Fibonacci F(n+2)
    for n_atoms in (3, 5, 6):
        po = obc.BlockadeChain(4, 8, 1e-9, n_atoms, 3.0, 7.0)
        states, gu, gw, i_g, i_z = chain_generators(n_atoms, 3.0, 7.0)
        assert states == obc.basis_states(n_atoms) and (i_g, i_z) == (po.i_g, po.i_z)
        assert np.array_equal(gu[0], 3.0 * po.h_drive) and np.array_equal(gu[1], -7.0 * po.h_number)
        assert gw[2][i_g, i_g] == 1 and gw[3][i_z, i_z] == 1 and np.count_nonzero(gw) == 2
        assert np.array_equal(gu[0], gu[0].conj().T)
