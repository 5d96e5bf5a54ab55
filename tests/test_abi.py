"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/parament.h
declares, the host-only entry points behave like the reference's, and the UNCHANGED reference wrapper binds it.
No compute call is made (no GPU here): Parament_create must fail loudly with code 30, never fall back."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "parament.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#define PARAMENT_API.*", "", text)
    return sorted(set(re.findall(r"PARAMENT_API[^;(]*?\b(\w+)\s*\(", text)))


def test_header_declares_reference_abi():
    names = declared_symbols()
    # the 20 symbols of the reference build + device_info (nm -D of the reference .so)
    for n in ["Parament_create", "Parament_destroy", "Parament_setHamiltonian", "Parament_equiprop",
              "Parament_setIterationCyclesManually", "Parament_automaticIterationCycles", "Parament_peekAtLastError",
              "Parament_errorMessage", "Parament_selectIterationCycles_fp32", "Parament_selectIterationCycles_fp64",
              "OneNorm", "OneNorm_fp64", "device_info"] + \
             [f"Parament_{x}_fp64" for x in ["create", "destroy", "setHamiltonian", "equiprop", "setIterationCyclesManually",
                                               "automaticIterationCycles", "peekAtLastError"]]:
        assert n in names, n


def test_library_exports_every_declared_symbol():
    from parament_b200._lib import EXPORTED, lib, library_path
    out = subprocess.run(["nm", "-D", "--defined-only", str(library_path())], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    for name in declared_symbols():
        assert name in exported, f"{name} declared in include/parament.h but not exported"
        assert getattr(lib, name) is not None
    assert set(EXPORTED) == set(declared_symbols())


def test_no_blas_or_torch_dependency():
    from parament_b200._lib import library_path
    out = subprocess.run(["ldd", str(library_path())], capture_output=True, text=True).stdout
    assert "cublas" not in out.lower() and "torch" not in out.lower() and "nccl" not in out.lower()


def test_error_messages_byte_identical():
    from parament_b200._lib import lib
    expect = {0: "Success", 10: "Memory allocation on the host failed.", 20: "Memory allocation on the device failed.",
              30: "Failed to initialize the cuBLAS library.", 50: "Invalid value.", 60: "Failed to execute cuBLAS function.",
              70: "Timestep too large", 80: "No hamiltonian set", 90: "Invalid quadrature selection.",
              1000: "Unknown error code", 12345: "Unknown error code"}   # parament.cpp:859-882
    for code, msg in expect.items():
        assert lib.Parament_errorMessage(code).decode() == msg


def test_iteration_cycle_tables():
    """Table semantics of parament.cpp:723-766 against the oracle package's independent restatement."""
    from oracle.reference_emulation import select_iteration_cycles
    from parament_b200._lib import lib
    xs = np.concatenate([np.logspace(-5, 1.2, 400), [0.032516793, 0.219062571, 0.368382, 12.65570085, 12.7584, 12.76, 13.0, 0.0]])
    for x in xs:
        assert lib.Parament_selectIterationCycles_fp32(1.0, float(x)) == select_iteration_cycles(1.0, float(x), "fp32")
        assert lib.Parament_selectIterationCycles_fp64(float(x), 1.0) == select_iteration_cycles(float(x), 1.0, "fp64")
    assert lib.Parament_selectIterationCycles_fp32(1.0, 0.2) == 5 and lib.Parament_selectIterationCycles_fp64(1.0, 0.2) == 11
    assert lib.Parament_selectIterationCycles_fp32(1.0, 13.0) == -1


def test_one_norm():
    from oracle.equiprop_oracle import one_norm
    from parament_b200._lib import lib
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 16):
        m = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        assert abs(lib.OneNorm_fp64(np.ascontiguousarray(m.ravel()), n) - one_norm(m)) < 1e-12
        assert abs(lib.OneNorm(np.ascontiguousarray(m.astype(np.complex64).ravel()), n) - one_norm(m.astype(np.complex64))) < 1e-5


def test_create_without_gpu_fails_loudly(gpu_count):
    if gpu_count > 0:
        pytest.skip("a GPU is present")
    import parament_b200
    with pytest.raises(RuntimeError, match="Error code 30"):
        parament_b200.Parament()
    h = ctypes.c_void_p()
    assert parament_b200._lib.lib.Parament_create_fp64(ctypes.byref(h)) == 30
    assert h.value is None
    assert parament_b200._lib.lib.Parament_destroy(None) == 0        # NULL handle is a no-op (parament.cpp:190)


def test_missing_library_fails_loudly(tmp_path):
    env = dict(os.environ, PARAMENT_LIB_DIR=str(tmp_path), PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", "import parament_b200"], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU or PyTorch fallback" in r.stderr


def _reference_wrapper_dir():
    for d in (os.path.join(ROOT, "oracle", "_ref", "pyparament"), "/root/reference/src/python/pyparament"):
        if os.path.isdir(os.path.join(d, "parament")):
            return d
    return None


def test_unchanged_reference_wrapper_binds_our_library():
    """paramentlib.py resolves ten symbols at import (paramentlib.py:57-72); a missing one is an AttributeError."""
    d = _reference_wrapper_dir()
    if d is None:
        pytest.skip("reference wrapper not staged (oracle/build_ref.sh)")
    from parament_b200._lib import DEFAULT_LIB_DIR
    env = dict(os.environ, PARAMENT_LIB_DIR=str(DEFAULT_LIB_DIR), PYTHONPATH=d)
    code = ("import numpy as np; np.float = float; import parament, parament.paramentlib as pl; "
            "print(pl.lib_path); "
            "import sys; "
            "ok = all(hasattr(pl.lib, s) for s in ['Parament_create','Parament_equiprop_fp64','device_info']); "
            "sys.exit(0 if ok else 3)")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert str(DEFAULT_LIB_DIR) in r.stdout
