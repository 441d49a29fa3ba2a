"""Generates `tests/golden/ref_*.npz`: outputs of the REFERENCE'S OWN SOURCE FILES
(/root/reference/code/quantum_optimal_control/{state_transfer,two_qubit}/propagator_vl.py, loaded
unmodified by path) executed on the torch-backed `jax` stand-in of `tests/golden/jax_shim/`.

    python tests/golden/make_reference_golden.py        # authoring container only (/root/reference)

What the fixtures pin: every formula, constant, index convention and quirk of the reference's Python
source, for the pulse map, the auxiliary amplitudes, the per-slice exponentials, the product tree, the
metrics, the cost AND its gradient (`jax.value_and_grad(target, argnums=(0,1,2))`, TQ:592 / notebook 01
cell 9, evaluated by reverse-mode autodiff THROUGH the reference's code).  What they do not pin: JAX's own
Pade order selection and XLA's summation orders (the stand-in always uses Pade-13 and torch matmul; ~1e-15).

The inputs are exactly those of `oracle.*.notebook0x_setup` / `oracle.ensemble.robust_cz_members`, so the
tests can rebuild the engine's problem from the seeds; the arrays are stored as well.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/code/quantum_optimal_control"
sys.path.insert(0, os.path.join(HERE, "jax_shim"))
sys.path.insert(0, ROOT)

import jax  # noqa: E402  (the stand-in)
import torch  # noqa: E402


def load_reference(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def npy(x):
    return x.detach().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def run_case(name, prop, a, b, c, n_exp=7, extra=None):
    ja, jb, jc = (jax.numpy.asarray(x) for x in (a, b, c))
    with torch.no_grad():
        wave = npy(prop.return_physical_amplitudes(ja, jb, jc))
        aux = npy(prop.return_auxiliary_amplitudes(ja, jb, jc))
        exps = prop.exponentials(ja, jb, jc)
        pick = np.linspace(0, exps.shape[0] - 1, n_exp).astype(int)
        exp_slices = npy(exps[pick.tolist()])
        vl = npy(prop.propagate(ja, jb, jc))
        infid, adiab = (float(x) for x in prop.metrics(ja, jb, jc))
    cost, (ga, gb, gc) = jax.value_and_grad(lambda x, y, z: prop.target(x, y, z), argnums=(0, 1, 2))(ja, jb, jc)
    out = dict(a=a, b=b, c=c, wave=wave, aux=aux, exp_index=pick, exp_slices=exp_slices, propagator=vl,
               infid=infid, adiab=adiab, cost=float(cost), ga=npy(ga), gb=npy(gb), gc=npy(gc),
               contraction_array=np.array(prop.contraction_array, dtype=bool), duration=float(prop.duration),
               generators=npy(prop.VL_Generators_ansatz))
    out.update(extra or {})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: cost {float(cost):.16f} infid {infid:.12f} adiab {adiab:.12f} |grad|max "
          f"{max(np.abs(npy(g)).max() for g in (ga, gb, gc)):.3e}")
    return out


def main():
    from oracle import ensemble, state_transfer as ost, two_qubit as otq
    ST = load_reference("state_transfer/propagator_vl.py", "ref_state_transfer")
    TQ = load_reference("two_qubit/propagator_vl.py", "ref_two_qubit")

    # ---- cfg 1: notebook 01 (the reference's stored 0.8022425853500806 is asserted in the tests)
    po, a, b, c = ost.notebook01_setup()
    rp = ST.PropagatorVL(po.input_dim, po.no_of_steps, po.delta_t, po.del_total, po.Delta_1, po.Rabi_1, po.Rabi_2,
                         po.Gamma_r, po.Gamma_i)
    run_case("ref_cfg1_state_transfer", rp, a, b, c)

    # ---- cfg 2: notebook 02 parameters (rng seed 0), and an odd-slice small grid
    def tq_ref(po):
        return TQ.PropagatorVL(po.input_dim, po.no_of_steps, po.padding, 50e6, po.delta_t, po.del_total, po.Vint,
                               po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
    po, a, b, c = otq.notebook02_setup()
    run_case("ref_cfg2_two_qubit", tq_ref(po), a, b, c)
    po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
    run_case("ref_small_two_qubit_odd", tq_ref(po), a, b, c)

    # ---- the reference's single_optimization_step (TQ:580-600) on the small grid: one GD step
    cost, na, nb, nc = TQ.single_optimization_step(tq_ref(po), *(jax.numpy.asarray(x) for x in (a, b, c)), lr=0.02)
    np.savez_compressed(os.path.join(HERE, "ref_small_two_qubit_step.npz"), a=a, b=b, c=c, lr=0.02, cost=float(cost),
                        new_a=npy(na), new_b=npy(nb), new_c=npy(nc))

    # ---- cfg 3 at its stated size (nt = 1888, pad = 56 -> M = 2000, N = 16): 4 members of the robust-CZ ensemble,
    #      each the reference class constructed with (del_total_s, Rabi * (1 + eps_s)); per-member cost / metrics /
    #      gradient.  Only scalars and gradients are stored (the propagators are 32x32 each).
    dl, rs = ensemble.robust_cz_noise(4096)
    members, (a, b, c) = ensemble.robust_cz_members(dl[:4], rs[:4], seed=0, n_gauss=16, nt=1888)
    costs, infids, adiabs, gas, gbs, gcs, vls = [], [], [], [], [], [], []
    ja, jb, jc = (jax.numpy.asarray(x) for x in (a, b, c))
    for po in members:
        rp = tq_ref(po)
        cost, (ga, gb, gc) = jax.value_and_grad(lambda x, y, z: rp.target(x, y, z), argnums=(0, 1, 2))(ja, jb, jc)
        with torch.no_grad():
            infid, adiab = (float(x) for x in rp.metrics(ja, jb, jc))
            vls.append(npy(rp.propagate(ja, jb, jc)))
        costs.append(float(cost)); infids.append(infid); adiabs.append(adiab)
        gas.append(npy(ga)); gbs.append(npy(gb)); gcs.append(npy(gc))
        print(f"ref_cfg3 member: cost {float(cost):.16f} infid {infid:.12f} adiab {adiab:.12f}")
    with torch.no_grad():
        wave = npy(tq_ref(members[0]).return_physical_amplitudes(ja, jb, jc))
    np.savez_compressed(os.path.join(HERE, "ref_cfg3_members.npz"), a=a, b=b, c=c, del_total=dl[:4], rabi_scale=rs[:4],
                        cost=np.array(costs), infid=np.array(infids), adiab=np.array(adiabs), ga=np.array(gas),
                        gb=np.array(gbs), gc=np.array(gcs), propagator=np.array(vls), wave=wave)


if __name__ == "__main__":
    main()
