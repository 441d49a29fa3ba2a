"""Stand-in for QuTiP: `from qutip import *` at TQ:8 pulls in nothing the PropagatorVL class uses."""
__all__ = []
