"""`parament` -- the package name of the reference's Python interface (pyparament, src/python/pyparament/parament/__init__.py:16),
provided by the B200-native library: code written against pyparament (`import parament; parament.Parament()`) runs on
parament_b200 without change after `pip install .` of this repository.

Same public names as the reference package: Parament, device_info, and the `constants` / `debug_functions` helpers."""
from parament_b200 import Parament, device_info, expm  # noqa: F401
from parament_b200 import constants  # noqa: F401
from . import debug_functions  # noqa: F401

__all__ = ["Parament", "device_info", "constants", "debug_functions"]
