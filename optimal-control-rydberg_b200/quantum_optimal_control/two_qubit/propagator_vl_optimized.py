"""Alias module: notebook 05 cell 24 of the reference imports
``two_qubit.propagator_vl_optimized.{PropagatorVLOpt, single_optimization_step_opt}``, also not in
the reference tree -- exactly the slot an accelerated backend fills."""
from .propagator_vl import PropagatorVL as PropagatorVLOpt  # noqa: F401
from .propagator_vl import single_optimization_step as single_optimization_step_opt  # noqa: F401

__all__ = ["PropagatorVLOpt", "single_optimization_step_opt"]
