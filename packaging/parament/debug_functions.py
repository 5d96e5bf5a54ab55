"""Counterpart of the reference's parament/debug_functions.py:22-31: the matrix exponential through the propagator."""
from parament_b200 import expm  # noqa: F401
