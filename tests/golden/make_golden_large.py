"""Golden fixtures of BASELINE.json configs 4 and 5 at their STATED sizes, from the dense oracle
(oracle/blockade_chain.py: dense 2n x 2n Van Loan exponentials per slice + torch autograd; slow on purpose, run once):

    python tests/golden/make_golden_large.py      # ~15 min on 8 cores, ~10 GB RAM

* cfg4_chain144_full.npz  -- 10-atom chain, blockade subspace n = 144 (VL 288), M = 4000 slices, T = 4 us, N = 16
                             Gaussians, rng seed 7; cost, metrics, 3N-gradient of the single sample.
* cfg5_chain377_members.npz -- 12-atom chain, n = 377 (VL 754), M = 256, T = 2 us, seed 100: cost, metrics and
                             3N-gradient of two members of the 512-sample noise ensemble (rng 1234), member by member.

There is no reference code for these models: the fixtures pin the ORACLE'S outputs, i.e. the dense
formulation in the reference's pattern, and say so in their `source` field.  A few slice exponentials are cross-checked
against scipy.linalg.expm while generating."""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import blockade_chain as obc  # noqa: E402

SOURCE = "oracle/blockade_chain.py::value_and_grad_columns (dense 2n x 2n expm per slice, torch autograd)"


def check_expm(po, a, b, c):
    with torch.no_grad():
        w = po.waves(*(torch.tensor(x) for x in (a, b, c)))
        M = po.no_of_steps
        for k in (0, M // 3, M - 1):
            e = po._dense_steps(w, k, k + 1)[0].numpy()
            n = po.dim
            ham = (po.rabi_scale * po.Rabi_max * float(w[k, 0])) * po.h_drive - \
                  (po.Delta_max * float(w[k, 1]) + po.del_offset) * po.h_number
            pen = np.zeros((n, n), complex)
            pen[po.i_g, po.i_g], pen[po.i_z, po.i_z] = 1.0 - k / M, k / M
            x = po.delta_t * np.block([[-1j * ham, pen], [np.zeros((n, n)), -1j * ham]])
            err = np.abs(e - sla.expm(x)).max()
            assert err < 1e-13, err


def cfg4():
    t0 = time.time()
    rng = np.random.default_rng(7)
    a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
    po = obc.BlockadeChain(16, 4000, 1e-9, 10, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    assert po.dim == 144
    check_expm(po, a, b, c)
    cost, infid, adiab, ga, gb, gc = po.value_and_grad_columns(a, b, c, chunk=100)
    np.savez_compressed(os.path.join(HERE, "cfg4_chain144_full.npz"), a=a, b=b, c=c, cost=cost, infid=infid, adiab=adiab,
                        ga=ga, gb=gb, gc=gc, source=SOURCE)
    print(f"cfg4: cost {cost:.15f} infid {infid:.12f} adiab {adiab:.12f}  ({time.time() - t0:.0f} s)", flush=True)


def cfg5(n_members=2):
    t0 = time.time()
    rng = np.random.default_rng(100)
    a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
    nrng = np.random.default_rng(1234)
    dl, rs = nrng.normal(0, 2 * np.pi * 0.1e6, 512), 1 + nrng.normal(0, 0.01, 512)
    rows = []
    for s in range(n_members):
        po = obc.BlockadeChain(16, 256, 2e-6 / 256, 12, 2 * np.pi * 2e6, 2 * np.pi * 4e6, del_offset=float(dl[s]),
                               rabi_scale=float(rs[s]))
        assert po.dim == 377
        if s == 0:
            check_expm(po, a, b, c)
        rows.append(po.value_and_grad_columns(a, b, c, chunk=32))
        print(f"cfg5 member {s}: cost {rows[-1][0]:.15f}  ({time.time() - t0:.0f} s)", flush=True)
    np.savez_compressed(os.path.join(HERE, "cfg5_chain377_members.npz"), a=a, b=b, c=c, del_total=dl[:n_members],
                        rabi_scale=rs[:n_members], cost=np.array([r[0] for r in rows]),
                        infid=np.array([r[1] for r in rows]), adiab=np.array([r[2] for r in rows]),
                        ga=np.array([r[3] for r in rows]), gb=np.array([r[4] for r in rows]),
                        gc=np.array([r[5] for r in rows]), source=SOURCE)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg5", "cfg4"]
    if "cfg5" in which:
        cfg5()
    if "cfg4" in which:
        cfg4()
