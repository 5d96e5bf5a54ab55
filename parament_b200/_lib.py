"""ctypes binding of libparament.so (the CUDA library is the product; there is no CPU fallback).

Mirrors /root/reference/src/python/pyparament/parament/paramentlib.py:17-72: same library name, same
PARAMENT_LIB_DIR override, same argtypes for the reference's entry points; adds the argtypes of the
additive entry points of include/parament.h section 2.
"""
from __future__ import annotations

import ctypes
import os
import pathlib

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
DEFAULT_LIB_DIR = _HERE / "lib"


def library_path() -> pathlib.Path:
    d = os.environ.get("PARAMENT_LIB_DIR")
    return (pathlib.Path(d) if d else DEFAULT_LIB_DIR) / "libparament.so"


def _load():
    path = library_path()
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `make -C {_HERE / 'csrc'}` (nvcc, sm_100a). "
            "parament_b200 has no CPU or PyTorch fallback.")
    return ctypes.cdll.LoadLibrary(str(path))


lib = _load()

ctx_p = ctypes.c_void_p
c64_p = np.ctypeslib.ndpointer(np.complex64, flags="C_CONTIGUOUS")
c128_p = np.ctypeslib.ndpointer(np.complex128, flags="C_CONTIGUOUS")
u32, u64, f64 = ctypes.c_uint, ctypes.c_ulonglong, ctypes.c_double

for suffix, cp in (("", c64_p), ("_fp64", c128_p)):
    getattr(lib, "Parament_create" + suffix).argtypes = [ctypes.POINTER(ctx_p)]
    getattr(lib, "Parament_destroy" + suffix).argtypes = [ctx_p]
    getattr(lib, "Parament_setHamiltonian" + suffix).argtypes = [ctx_p, cp, cp, u32, u32, ctypes.c_bool, ctypes.c_int]
    getattr(lib, "Parament_equiprop" + suffix).argtypes = [ctx_p, cp, f64, u32, u32, cp]
    getattr(lib, "Parament_equipropBatch" + suffix).argtypes = [ctx_p, cp, f64, u32, u32, u32, cp]
    getattr(lib, "Parament_equipropDevice" + suffix).argtypes = [ctx_p, ctypes.c_void_p, f64, u32, u32, u32, ctypes.c_void_p, ctypes.c_void_p]
    getattr(lib, "Parament_equipropSlice" + suffix).argtypes = [ctx_p, cp, f64, u32, u32, u64, u64, cp]
    getattr(lib, "Parament_equipropSliceToDevice" + suffix).argtypes = [ctx_p, cp, f64, u32, u32, u64, u64, ctypes.c_void_p]
    getattr(lib, "Parament_combine" + suffix).argtypes = [ctx_p, cp, u32, cp]
    getattr(lib, "Parament_setIterationCyclesManually" + suffix).argtypes = [ctx_p, u32]
    getattr(lib, "Parament_automaticIterationCycles" + suffix).argtypes = [ctx_p]
    getattr(lib, "Parament_peekAtLastError" + suffix).argtypes = [ctx_p]

lib.Parament_errorMessage.argtypes = [ctypes.c_int]
lib.Parament_errorMessage.restype = ctypes.c_char_p
lib.Parament_getLastError.argtypes = [ctx_p]
lib.Parament_selectIterationCycles_fp32.argtypes = [f64, f64]
lib.Parament_selectIterationCycles_fp64.argtypes = [f64, f64]
lib.OneNorm.argtypes = [c64_p, u32]
lib.OneNorm.restype = f64
lib.OneNorm_fp64.argtypes = [c128_p, u32]
lib.OneNorm_fp64.restype = f64
lib.device_info.argtypes = []
lib.device_info.restype = None
lib.Parament_combineDevice.argtypes = [ctx_p, ctypes.c_void_p, u32, ctypes.c_void_p, ctypes.c_void_p]
lib.Parament_lastStat.argtypes = [ctx_p, ctypes.c_int]
lib.Parament_lastStat.restype = f64
lib.Parament_setDevice.argtypes = [ctx_p, ctypes.c_int]
lib.Parament_setDevices.argtypes = [ctx_p, ctypes.c_int]
lib.Parament_setDeviceList.argtypes = [ctx_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
lib.Parament_measurePeak.argtypes = [ctypes.c_int]
lib.Parament_measurePeak.restype = f64
lib.Parament_version.argtypes = []
lib.Parament_version.restype = ctypes.c_char_p

EXPORTED = [
    # section 1: the reference's ABI (nm -D of the reference build)
    "Parament_create", "Parament_destroy", "Parament_setHamiltonian", "Parament_equiprop",
    "Parament_setIterationCyclesManually", "Parament_automaticIterationCycles", "Parament_peekAtLastError",
    "Parament_create_fp64", "Parament_destroy_fp64", "Parament_setHamiltonian_fp64", "Parament_equiprop_fp64",
    "Parament_setIterationCyclesManually_fp64", "Parament_automaticIterationCycles_fp64",
    "Parament_peekAtLastError_fp64", "Parament_errorMessage", "Parament_selectIterationCycles_fp32",
    "Parament_selectIterationCycles_fp64", "OneNorm", "OneNorm_fp64", "device_info", "Parament_getLastError",
    # section 2: additive
    "Parament_equipropBatch", "Parament_equipropBatch_fp64", "Parament_equipropDevice",
    "Parament_equipropDevice_fp64", "Parament_equipropSlice", "Parament_equipropSlice_fp64", "Parament_equipropSliceToDevice",
    "Parament_equipropSliceToDevice_fp64", "Parament_combine",
    "Parament_combine_fp64", "Parament_combineDevice", "Parament_lastStat", "Parament_setDevice", "Parament_setDevices",
    "Parament_setDeviceList", "Parament_measurePeak", "Parament_version",
]
