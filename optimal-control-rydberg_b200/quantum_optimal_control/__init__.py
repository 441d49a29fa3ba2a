"""quantum_optimal_control -- drop-in mirror of the reference package's hot path
(reference: code/quantum_optimal_control/__init__.py:5-11), backed by the B200 engine ``qocvl``.

Only the Van Loan cost / gradient path is provided: ``state_transfer.propagator_vl`` and
``two_qubit.propagator_vl`` (+ the two alias modules the reference notebooks import).  The
plotting toolkits and the QuTiP validation scripts are out of scope and are NOT imported here,
so the package imports without matplotlib / qutip / arc.
"""
__version__ = "0.1.0"   # reference __init__.py:5

from . import state_transfer, two_qubit, single_qubit, blockade_chain  # noqa: E402,F401
