"""parament_b200 -- B200-native Parament_equiprop (CUDA library + the host-side mirror of the reference interface).

Importing this package loads parament_b200/lib/libparament.so and fails loudly if it is missing.
"""
from .parament import Parament, device_info, expm  # noqa: F401
from . import constants  # noqa: F401
