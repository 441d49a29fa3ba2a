from . import propagator_vl  # noqa: F401
