"""Alias of ``qutip_five_level`` under the reference's other module name (``two_qubit/qutip_simulation.py``: the same
three functions, QS:4-103)."""
from .qutip_five_level import (QTHamiltonian, Mesolve_5lvl_t, gen_n_level_atom_basis, tensor,  # noqa: F401
                               collapse_operators, liouvillian_pieces)
