from . import propagator_vl, propagator_vl_jax, propagator_vl_optimized  # noqa: F401
