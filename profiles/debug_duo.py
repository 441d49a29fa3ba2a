import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/optimal-control-rydberg_b200")
import numpy as np, qocvl
from oracle import two_qubit as otq, ensemble
from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL
po, a, b, c = otq.notebook02_setup(seed=3, n_gauss=6, nt=77)
args = (6, 77, po.padding, 50e6, po.delta_t, 0.0, po.Vint, po.Delta_i_val, po.Rabi_i_val, po.Rabi_r_val, otq.GAMMAS_CZ)
dl, rs = ensemble.robust_cz_noise(5)
with qocvl.options(debug_sync=1):
    small = PropagatorVL(*args)
small.set_ensemble(del_total=dl, rabi_scale=rs)
try:
    out = small.cost_and_grad_packed(small._pack(a, b, c)).cpu().numpy()
    print("ok", out[:3])
except Exception as e:
    print("FAILED:", e)
