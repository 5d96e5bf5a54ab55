"""Generate tests/golden/ref_cuda.npz: outputs of the REFERENCE's own CUDA build for the seeded cases of
cases.py.  Runs on a GPU box only:

    oracle/build_ref.sh                        # in the dev container: compiles /root/reference -> oracle/_ref/
    gpurun -- python tests/golden/make_ref_cuda_golden.py gpurun_out/ref_cuda.npz
    cp gpurun_out/ref_cuda.npz tests/golden/

Calls oracle/_ref/libparament.so directly through ctypes with the signatures of the reference header
(parament.h:155-400); no file of the reference tree is read at run time.  Only configurations on which the
reference is well defined are run (odd MMAX, Magnus with <= 3 controls and all amplitude arrays given).
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import extended_cases, reference_test_cases  # noqa: E402

QUAD = {"none": 0, "midpoint": 0x01000000, "simpson": 0x02000000}


def run_reference(lib, case):
    fp64 = case["precision"] == "fp64"
    sfx = "_fp64" if fp64 else ""
    ct = np.complex128 if fp64 else np.complex64
    h = ctypes.c_void_p()
    assert getattr(lib, "Parament_create" + sfx)(ctypes.byref(h)) == 0
    H0 = np.ascontiguousarray(case["H0"].astype(ct).ravel())
    H1 = np.ascontiguousarray(case["H1"].astype(ct).ravel())
    n, A = case["H0"].shape[0], case["H1"].shape[0]
    carr = np.ascontiguousarray(case["carr"].astype(ct).ravel())
    Au, pts = case["carr"].shape
    out = np.zeros(n * n, dtype=ct)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ec = getattr(lib, "Parament_setHamiltonian" + sfx)(h, vp(H0), vp(H1), ctypes.c_uint(n), ctypes.c_uint(A),
                                                       ctypes.c_bool(case["use_magnus"]), ctypes.c_int(QUAD[case["quadrature"]]))
    assert ec == 0, ec
    if case.get("mmax"):
        getattr(lib, "Parament_setIterationCyclesManually" + sfx)(h, ctypes.c_uint(case["mmax"]))
    ec = getattr(lib, "Parament_equiprop" + sfx)(h, vp(carr), ctypes.c_double(case["dt"]), ctypes.c_uint(pts), ctypes.c_uint(Au), vp(out))
    getattr(lib, "Parament_destroy" + sfx)(h)
    if ec != 0:
        return None
    return out.reshape(n, n)


def reference_defined(case):
    n_steps = case["carr"].shape[1]
    q, mag = case["quadrature"], case["use_magnus"]
    N = (n_steps - 1) // 2 if (mag or q == "simpson") else (n_steps - 1 if q == "midpoint" else n_steps)
    if N < 1:
        return False          # reference returns stale memory (SURVEY A-7)
    if mag and (case["H1"].shape[0] > 3 or case["carr"].shape[0] != case["H1"].shape[0]):
        return False          # non-injective slot map / mixed layouts (SURVEY A-3, A-4)
    if case.get("mmax") and case["mmax"] % 2 == 0:
        return False
    return True


def main(out_path):
    root = os.path.dirname(os.path.dirname(HERE))
    lib = ctypes.cdll.LoadLibrary(os.path.join(root, "oracle", "_ref", "libparament.so"))
    store = {}
    for case in reference_test_cases() + extended_cases():
        if not reference_defined(case):
            continue
        U = run_reference(lib, case)
        if U is None:
            print("reference returned an error for", case["name"])
            continue
        store[case["name"]] = U
        print(f"{case['name']:40s} |U|_F = {np.linalg.norm(U):.6f}", flush=True)
    np.savez_compressed(out_path, **store)
    print("wrote", out_path, len(store), "cases")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_cuda.npz"))
