"""Alias module: notebooks 02 (cell 2) and 04 (cell 3) of the reference import
``two_qubit.propagator_vl_jax.PropagatorVLJAX``, a module that is not in the reference tree;
it is the class of two_qubit/propagator_vl.py."""
from .propagator_vl import PropagatorVL as PropagatorVLJAX, single_optimization_step  # noqa: F401

__all__ = ["PropagatorVLJAX", "single_optimization_step"]
