"""CPU tests: the oracle against every known answer the reference holds for the path, against the outputs of
the reference's own CUDA build (generated on a B200, tests/golden/ref_cuda.npz) and against the numpy
emulation of the reference's algorithm."""
import os

import numpy as np
import pytest
import scipy.linalg

from cases import extended_cases, reference_test_cases
from oracle.equiprop_oracle import (effective_dt, effective_steps, equiprop_oracle, hnorm, ordered_product,
                                    rel_frobenius)
from oracle.reference_emulation import reference_equiprop_emulated, select_iteration_cycles

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF_TESTS = np.load(os.path.join(GOLD, "reference_tests.npz"))
REF_CUDA = np.load(os.path.join(GOLD, "ref_cuda.npz")) if os.path.exists(os.path.join(GOLD, "ref_cuda.npz")) else None


def run_oracle(case, **kw):
    return equiprop_oracle(case["H0"], case["H1"], case["carr"], case["dt"], case["quadrature"], case["use_magnus"],
                           case["precision"], **kw)


@pytest.mark.parametrize("case", reference_test_cases(), ids=lambda c: c["name"])
def test_oracle_meets_reference_test_thresholds(case):
    """The oracle must itself pass the reference's acceptance thresholds (test_numerics.py:46,66,86,117).
    Inputs are rounded to the context precision first, so fp32 cases carry the input rounding only."""
    U = run_oracle(case)
    exp = REF_TESTS[case["name"]]
    if case["kind"] == "sumabs":
        # the oracle evaluates the same scipy expression on precision-rounded inputs
        err = np.sum(np.abs(U - exp))
        tol = case["threshold"] if case["precision"] == "fp64" else 4 * case["threshold"]
        assert err < tol
    else:
        assert np.linalg.norm(U - exp) < case["threshold"]


def test_docstring_known_answer():
    case = [c for c in reference_test_cases() if c["name"] == "docstring_kat"][0]
    U = run_oracle(case)
    assert np.allclose(U, REF_TESTS["docstring_kat"], atol=5e-8)
    assert np.allclose(U, np.diag([np.exp(-1j), np.exp(1j)]), atol=1e-7)


@pytest.mark.parametrize("case", [c for c in extended_cases() if c.get("mmax") is None], ids=lambda c: c["name"])
def test_oracle_vs_reference_cuda(case):
    """Oracle vs the reference's CUDA build executed on a B200 (short N: reference error is ~1e-8 N in
    fp32 and ~3e-17 N in fp64)."""
    if REF_CUDA is None or case["name"] not in REF_CUDA.files:
        pytest.skip("no reference-CUDA vector for this case (reference undefined there or file not generated)")
    U = run_oracle(case)
    N, _ = effective_steps(case["carr"].shape[1], case["quadrature"], case["use_magnus"])
    tol = (2e-6 + 3e-8 * N) if case["precision"] == "fp32" else (1e-13 + 1e-15 * N)
    assert rel_frobenius(REF_CUDA[case["name"]], U) < tol


@pytest.mark.parametrize("case", [c for c in extended_cases() if c["H0"].shape[0] <= 16 and c["carr"].shape[1] <= 210
                                  and (not c["use_magnus"] or c["H1"].shape[0] <= 3)][:24], ids=lambda c: c["name"])
def test_oracle_vs_reference_emulation(case):
    """Closed-form quadrature / Magnus semantics of the oracle vs a step-by-step numpy emulation of
    parament.cpp + control_expansion.cu in float64 (agreement ~1e-14 pins the semantics)."""
    if case["carr"].shape[0] != case["H1"].shape[0] and case["use_magnus"]:
        pytest.skip("reference mixes slot layouts")
    N, _ = effective_steps(case["carr"].shape[1], case["quadrature"], case["use_magnus"])
    if N < 1:
        pytest.skip("reference undefined for zero effective steps")
    U = equiprop_oracle(case["H0"], case["H1"], case["carr"], case["dt"], case["quadrature"], case["use_magnus"], "fp64")
    h = effective_dt(case["dt"], case["quadrature"], case["use_magnus"])
    m = case.get("mmax") or select_iteration_cycles(hnorm(case["H0"].astype(np.complex128), case["H1"].astype(np.complex128)), h, "fp64")
    if m % 2 == 0 or m < 3:
        pytest.skip("reference recurrence undefined for this MMAX")
    V = reference_equiprop_emulated(case["H0"], case["H1"], case["carr"], case["dt"], case["quadrature"], case["use_magnus"],
                                    "fp64", mmax=m if case.get("mmax") else None)
    tol = 1e-12 if not case.get("mmax") else 1e-6
    assert rel_frobenius(V, U) < tol


def test_effective_step_rules():
    assert effective_steps(10, "none") == (10, 1)
    assert effective_steps(10, "midpoint") == (9, 1)
    assert effective_steps(10, "simpson") == (4, 2)      # even P: trailing point dropped (parament.cpp:825)
    assert effective_steps(11, "simpson") == (5, 2)
    assert effective_steps(11, "simpson", True) == (5, 2)
    assert effective_steps(1, "midpoint") == (0, 1)
    assert effective_steps(2, "simpson") == (0, 2)
    assert effective_dt(0.1, "simpson") == 0.2 and effective_dt(0.1, "midpoint") == 0.1


def test_zero_steps_is_identity():
    H0 = np.diag([1.0, -1.0])
    U = equiprop_oracle(H0, H0[None], np.zeros((1, 1)), 0.1, "midpoint")
    assert np.array_equal(U, np.eye(2))


def test_ordered_product_order():
    rng = np.random.default_rng(0)
    U = rng.standard_normal((7, 3, 3)) + 1j * rng.standard_normal((7, 3, 3))
    ref = np.eye(3, dtype=complex)
    for j in range(7):
        ref = U[j] @ ref
    assert np.allclose(ordered_product(U), ref)


def test_parallel_slices_match_serial():
    from workloads import make_workload
    w = make_workload("C2", pts=4001)
    a = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision, workers=1)
    b = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision, workers=3, block=97)
    assert rel_frobenius(a, b) < 1e-13


def test_oracle_unitarity_and_magnus_order():
    """Magnus must converge at 4th order, Simpson/midpoint at 2nd: halving dt on a smooth
    pulse reduces the error by ~16x / ~4x."""
    rng = np.random.default_rng(5)
    n = 4
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H0 = (g + g.conj().T) / 4
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H1 = (g + g.conj().T) / 4
    T = 2.0

    def prop(P, quad, mag):
        t = np.linspace(0, T, P)
        c = np.sin(3 * t) + 0.3 * t
        return equiprop_oracle(H0, H1[None], c[None], T / (P - 1), quad, mag, "fp64")

    truth = prop(20481, "simpson", True)
    assert np.linalg.norm(truth.conj().T @ truth - np.eye(n)) < 1e-12
    for quad, mag, order in (("midpoint", False, 2), ("simpson", False, 2), ("simpson", True, 4)):
        e1 = np.linalg.norm(prop(41, quad, mag) - truth)
        e2 = np.linalg.norm(prop(81, quad, mag) - truth)
        assert 0.6 * 2 ** order < e1 / e2 < 1.6 * 2 ** order, (quad, mag, e1 / e2)
