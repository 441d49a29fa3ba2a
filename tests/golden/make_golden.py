"""Generates the committed golden fixtures from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference itself (JAX) cannot run here, so the
fixtures pin the ORACLE's outputs; the one number the reference stores (notebook 01 cell 5) is
asserted separately in tests/test_oracle.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import autograd_ref, state_transfer, two_qubit  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def dump(name, prop, a, b, c, extra=None):
    cost, infid, adiab, ga, gb, gc = autograd_ref.value_and_grad(prop, a, b, c)
    wave = prop.return_physical_amplitudes(a, b, c)
    aux = prop.return_auxiliary_amplitudes(a, b, c)
    exps = prop.exponentials(a, b, c)
    vl = prop.propagate(a, b, c)
    pick = np.linspace(0, exps.shape[0] - 1, 7).astype(int)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), a=a, b=b, c=c, cost_np=prop.target(a, b, c),
        metrics_np=np.array(prop.metrics(a, b, c)), cost=cost, infid=infid, adiab=adiab, ga=ga, gb=gb, gc=gc,
        wave=wave, aux=aux, exp_index=pick, exp_slices=exps[pick], propagator=vl, **(extra or {}))
    print(name, cost, infid, adiab)


if __name__ == "__main__":
    p, a, b, c = state_transfer.notebook01_setup()
    dump("cfg1_state_transfer", p, a, b, c)
    p, a, b, c = two_qubit.notebook02_setup()
    dump("cfg2_two_qubit", p, a, b, c)
    # odd slice count exercises the "fold the leftover" branch at several tree levels
    p, a, b, c = two_qubit.notebook02_setup(seed=3, n_gauss=6, nt=77)
    dump("small_two_qubit_odd", p, a, b, c)
