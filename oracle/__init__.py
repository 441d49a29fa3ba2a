"""CPU oracle for the Van Loan cost + gradient path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy (forward) and torch-CPU-autograd (gradient) restatement
of the reference's two ``PropagatorVL`` classes

  * /root/reference/code/quantum_optimal_control/state_transfer/propagator_vl.py  ("ST")
  * /root/reference/code/quantum_optimal_control/two_qubit/propagator_vl.py      ("TQ")

PINNING.  The reference is pure Python on JAX; jax / jaxlib cannot be installed in this image (no network), but its two
source files need only ~45 ``jax.numpy`` names, ``expm``, ``vmap``, ``jit`` and ``value_and_grad``.  ``tests/golden/jax_shim``
provides exactly those on torch-CPU complex128, so ``tests/golden/make_reference_golden.py`` EXECUTES THE REFERENCE'S OWN
FILES UNMODIFIED (loaded by path from /root/reference) and stores their outputs as ``tests/golden/ref_*.npz``: pulse map,
auxiliary amplitudes, slice exponentials, propagator, metrics, cost and the full ``jax.value_and_grad`` gradient for
notebook 01 (cfg 1), notebook 02 (cfg 2), an odd-slice CZ grid, one ``single_optimization_step`` and four members of cfg 3
at its stated size.  ``tests/test_reference_golden.py`` holds this oracle (CPU) and the CUDA engine (GPU) to those
fixtures; the one number the reference itself stores (notebook 01 cell 5, ``Initial Figure of Merit:
0.8022425853500806``) is reproduced as well (tests/test_oracle.py).  What the fixtures cannot pin is JAX's own Pade-order
selection and XLA's summation orders (~1e-15).  Beside that the three independent formulations in this package (SciPy
Pade expm + product tree, torch Taylor expm + autograd, the state-vector Taylor / adjoint model of vector_model.py) agree
with each other.

Parity unpinned by the reference, and said so where used: the blockade-chain model (BASELINE configs 4-5 have no reference
source: oracle/blockade_chain.py DEFINES them per SURVEY 8(d)), the noise-ensemble semantic (oracle/ensemble.py: the
reference class re-instantiated per sample) and the Lindblad validation (no stored mesolve output; two independent
integrators instead).  oracle/norm_bounds.py restates the engine's Taylor-degree rule (no reference counterpart).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (qocvl, quantum_optimal_control) never does.
"""
