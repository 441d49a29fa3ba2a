"""CPU oracle for the Van Loan cost + gradient path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy (forward) and torch-CPU-autograd (gradient) restatement
of the reference's two ``PropagatorVL`` classes

  * /root/reference/code/quantum_optimal_control/state_transfer/propagator_vl.py  ("ST")
  * /root/reference/code/quantum_optimal_control/two_qubit/propagator_vl.py      ("TQ")

The reference itself is JAX and cannot be imported in this image (jax, jaxlib,
matplotlib, qutip, arc are absent and there is no network), so parity is pinned by

  * the only known answer the reference stores: notebook 01 cell 5,
    ``Initial Figure of Merit: 0.8022425853500806`` (tests/test_oracle.py), and
  * agreement of three independent formulations of the same numbers
    (SciPy Pade expm + product tree, torch Taylor expm + autograd, and the
    state-vector Taylor/adjoint model in oracle/vector_model.py).

Gradient parity, the CZ model (notebook 02 has no stored outputs) and the
ensemble / PXP extrapolations are therefore "parity unpinned by the reference":
they are pinned only by this oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (qocvl, quantum_optimal_control) never does.
"""
