"""Generates the committed golden fixtures from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference itself (JAX) cannot run here, so the
fixtures pin the ORACLE's outputs; the one number the reference stores (notebook 01 cell 5) is
asserted separately in tests/test_oracle.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import autograd_ref, blockade_chain, state_transfer, two_qubit  # noqa: E402

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


def chain_ensemble_case(n_atoms=8, n_gauss=5, nt=48, seed=21, n_samples=3):
    """Blockade-chain ensemble (the model of BASELINE.json configs 4-5 at reduced size, n = 55): inputs and the
    member-by-member dense-oracle mean of cost and gradient."""
    po, a, b, c = blockade_chain.chain_setup(n_atoms=n_atoms, n_gauss=n_gauss, nt=nt, seed=seed)
    rng = np.random.default_rng(seed + 1)
    dl, rs = rng.normal(0, 2 * np.pi * 0.1e6, n_samples), 1 + rng.normal(0, 0.01, n_samples)
    acc = None
    for d, s in zip(dl, rs):
        member, *_ = blockade_chain.chain_setup(n_atoms=n_atoms, n_gauss=n_gauss, nt=nt, seed=seed,
                                                del_offset=float(d), rabi_scale=float(s))
        vals = [np.asarray(x, float) for x in member.value_and_grad(a, b, c)]
        acc = vals if acc is None else [x + y for x, y in zip(acc, vals)]
    mean = [x / n_samples for x in acc]
    return po, a, b, c, dl, rs, mean


def dump_chain(name):
    import torch
    po, a, b, c, dl, rs, mean = chain_ensemble_case()
    wave = po.waves(*(torch.tensor(x) for x in (a, b, c))).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), a=a, b=b, c=c, del_total=dl, rabi_scale=rs, cost=mean[0],
                        infid=mean[1], adiab=mean[2], ga=mean[3], gb=mean[4], gc=mean[5], wave=wave)
    print(name, mean[0], mean[1], mean[2])


if __name__ == "__main__":
    dump_chain("chain8_ensemble")
    p, a, b, c = state_transfer.notebook01_setup()
    dump("cfg1_state_transfer", p, a, b, c)
    p, a, b, c = two_qubit.notebook02_setup()
    dump("cfg2_two_qubit", p, a, b, c)
    # odd slice count exercises the "fold the leftover" branch at several tree levels
    p, a, b, c = two_qubit.notebook02_setup(seed=3, n_gauss=6, nt=77)
    dump("small_two_qubit_odd", p, a, b, c)
