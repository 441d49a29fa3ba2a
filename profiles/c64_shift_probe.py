import os, sys
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")): sys.path.insert(0, p)
import numpy as np, torch, qocvl
from oracle import two_qubit as otq, ensemble
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
po, a, b, c = otq.notebook02_setup(seed=8, n_gauss=6, nt=77)
dl, rs = ensemble.robust_cz_noise(9)
for mode, extra in ((2, {}), (3, {}), (3, {'taylor_tol_exp': 40}), (2, {'taylor_tol_exp': 40})):
    with qocvl.options(duo_roleq=mode, **extra):
        prop = PropagatorVL(6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
        prop.set_ensemble(del_total=dl, rabi_scale=rs)
        full = prop.cost_and_grad_packed(prop._pack(a,b,c)).cpu().numpy().copy()
        hist_128, *_ = prop.optimize(a, b, c, num_iters=12, lr=0.02)
        prop.set_precision("complex64")
        out = prop.cost_and_grad_packed(prop._pack(a,b,c)).cpu().numpy().copy()
        hist_g, *_ = prop.optimize(a, b, c, num_iters=12, lr=0.02)
        print(mode, extra, 'diffs', np.abs(out[:3]-full[:3]), 'single eval diff cost %.2e grad %.2e'%(abs(out[0]-full[0]), np.abs(out[3:]-full[3:]).max()/np.abs(full[3:]).max()), 'hist diff', (hist_g-hist_128).abs().cpu().numpy().round(7), prop._plan.role_degree_means())
