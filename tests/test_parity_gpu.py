"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C-ABI of libparament.so
(parament_b200.Parament is a thin ctypes mirror of the reference's wrapper) and is compared with
  * the frozen known answers of the reference's own tests (tests/golden/reference_tests.npz) at the
    reference's own thresholds,
  * outputs of the reference's CUDA build on a B200 (tests/golden/ref_cuda.npz),
  * the float64 scipy.linalg.expm product oracle at the north_star tolerances:
        relative Frobenius 1e-5 (complex64) / 1e-12 (complex128),
  * at BASELINE.json's full sizes: precomputed oracle propagators (tests/golden/full_*.npz) and
    size-independent properties (unitarity, slice composition, time reversal).
"""
import ctypes
import os
import threading

import numpy as np
import pytest

from cases import extended_cases, reference_test_cases
from oracle.equiprop_oracle import effective_steps, equiprop_oracle, rel_frobenius
from workloads import make_workload

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 1e-5, "fp64": 1e-12}          # north_star tolerances (relative Frobenius norm)


def _load(name):
    p = os.path.join(GOLD, name)
    return np.load(p) if os.path.exists(p) else None


REF_TESTS = _load("reference_tests.npz")
REF_CUDA = _load("ref_cuda.npz")


@pytest.fixture(scope="module")
def pb():
    import parament_b200
    return parament_b200


def run_case(pb, case):
    with pb.Parament(case["precision"]) as ctx:
        ctx.set_hamiltonian(case["H0"], *case["H1"], use_magnus=case["use_magnus"], quadrature_mode=case["quadrature"])
        if case.get("mmax"):
            ctx.set_iteration_cycles(case["mmax"])
        U = ctx.equiprop(case["dt"], *case["carr"])
        fam = ctx.stat(5)
    assert U.dtype == (np.complex64 if case["precision"] == "fp32" else np.complex128)
    return U, fam


# ---------------------------------------------------------------------------------------------------------
# the reference's own acceptance tests, same inputs, same thresholds
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", reference_test_cases(), ids=lambda c: c["name"])
def test_reference_known_answers(pb, case):
    U, _ = run_case(pb, case)
    exp = REF_TESTS[case["name"]]
    if case["kind"] == "sumabs":
        assert np.sum(np.abs(U - exp)) < case["threshold"]          # test_numerics.py:46,66,86
    else:
        assert np.linalg.norm(U - exp) < case["threshold"]          # test_numerics.py:117


def test_debug_expm(pb):
    """parament.debug_functions.expm (debug_functions.py:22-31): non-Hermitian generator H0 = i m."""
    import scipy.linalg
    rng = np.random.default_rng(1)
    m = 0.2 * (rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))
    assert np.sum(np.abs(pb.expm(m) - scipy.linalg.expm(m))) < np.finfo(np.float32).eps * 16 * 4


# ---------------------------------------------------------------------------------------------------------
# every quadrature / Magnus / ragged size / dimension family against the oracle and the reference CUDA build
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", extended_cases(), ids=lambda c: c["name"])
def test_extended_cases_vs_oracle(pb, case):
    U, fam = run_case(pb, case)
    n = case["H0"].shape[0]
    assert fam == (1 if n <= 16 else (2 if n <= 64 else 3))
    if case.get("mmax"):
        # a forced degree truncates the series: compare with the oracle only when the degree is sufficient
        from oracle.reference_emulation import select_iteration_cycles
        from oracle.equiprop_oracle import effective_dt, hnorm
        h = effective_dt(case["dt"], case["quadrature"], case["use_magnus"])
        need = select_iteration_cycles(hnorm(case["H0"], case["H1"]), h, case["precision"])
        if case["mmax"] < need:
            pytest.skip("forced degree below the table value")
    Uo = equiprop_oracle(case["H0"], case["H1"], case["carr"], case["dt"], case["quadrature"], case["use_magnus"], case["precision"])
    assert rel_frobenius(U, Uo) < TOL[case["precision"]]


@pytest.mark.parametrize("case", [c for c in extended_cases() if c["name"].startswith("fp64")], ids=lambda c: c["name"])
def test_extended_cases_clenshaw_path(pb, case, monkeypatch):
    """The same cases with the Horner-in-Y^2 evaluation disabled: the plain Clenshaw recurrence of the reference
    (parament.cpp:569-652) must give the same propagators (it is also what runs for Hnorm*h > 1)."""
    monkeypatch.setenv("PARAMENT_SERIES", "clenshaw")
    with pb.Parament(case["precision"]) as ctx:
        ctx.set_hamiltonian(case["H0"], *case["H1"], use_magnus=case["use_magnus"], quadrature_mode=case["quadrature"])
        if case.get("mmax"):
            ctx.set_iteration_cycles(case["mmax"])
        U = ctx.equiprop(case["dt"], *case["carr"])
        assert ctx.stat(9) == 0
    monkeypatch.delenv("PARAMENT_SERIES")
    with pb.Parament(case["precision"]) as ctx:
        ctx.set_hamiltonian(case["H0"], *case["H1"], use_magnus=case["use_magnus"], quadrature_mode=case["quadrature"])
        if case.get("mmax"):
            ctx.set_iteration_cycles(case["mmax"])
        V = ctx.equiprop(case["dt"], *case["carr"])
    assert rel_frobenius(U, V) < 1e-13


@pytest.mark.parametrize("n,A,quad,mag,complex_amps,prec", [
    (2, 1, "none", False, False, "fp32"), (5, 2, "midpoint", False, False, "fp32"), (8, 2, "simpson", False, True, "fp32"),
    (9, 1, "simpson", False, False, "fp32"), (16, 2, "simpson", False, False, "fp32"), (16, 3, "none", False, True, "fp32"),
    (16, 2, "simpson", True, False, "fp32"), (13, 1, "midpoint", False, True, "fp32"), (7, 2, "simpson", True, True, "fp32"),
    (24, 2, "none", False, True, "fp32"), (64, 2, "simpson", False, False, "fp32"), (100, 1, "none", False, False, "fp32"),
    (40, 2, "midpoint", False, True, "fp64"), (64, 3, "simpson", True, False, "fp64"), (96, 2, "none", False, False, "fp64")])
def test_degree8_three_product_path(pb, n, A, quad, mag, complex_amps, prec, monkeypatch):
    """Table degrees 6..8 are evaluated as one degree-8 polynomial in three matrix products (api.cu solve_degree8;
    k1_warp.cu for complex64 contexts of dim <= 16, k4_onchip.cu / k4_gemm.cu build_program above that).  It must be
    active, agree with the oracle, and agree with the Horner-in-Y^2 evaluation of the table degree -- for Hermitian and
    non-Hermitian generators (complex amplitudes), every quadrature, and Magnus terms beyond the software-pipelined ones.
    The series is built for the reference's norm bound here (PARAMENT_NORM=reference, read at Parament_create) so that the
    degree under test is the table's at every dimension; the spectral bound of dim > 16 is covered in test_round2_gpu.py."""
    monkeypatch.setenv("PARAMENT_NORM", "reference")
    monkeypatch.setenv("PARAMENT_C64_MATH", "f64")     # the FP64 kernels are under test (the TF32 path of dim <= 8: test_round2_gpu.py)
    rng = np.random.default_rng(100 * n + A)
    herm = lambda: (lambda g: (g + g.conj().T) / 2)(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    norm1 = lambda m: m / np.max(np.sum(np.abs(m), axis=1))
    ctype = np.complex64 if prec == "fp32" else np.complex128
    H0 = (0.6 * norm1(herm())).astype(ctype)
    H1 = [(0.4 / A * norm1(herm())).astype(ctype) for _ in range(A)]
    pts = 301 if n <= 64 else 41
    carr = rng.uniform(-1, 1, (A, pts)) + (1j * rng.uniform(-1, 1, (A, pts)) if complex_amps else 0)
    carr = carr.astype(ctype)
    x = 0.44 if prec == "fp32" else 0.04                              # Hnorm * h: table degree 7
    dt = x if quad in ("none", "midpoint") else x / 2
    tol = 2e-6 if prec == "fp32" else 1e-13

    def run():
        with pb.Parament(prec) as ctx:
            ctx.set_hamiltonian(H0, *H1, use_magnus=mag, quadrature_mode=quad)
            U = ctx.equiprop(dt, *carr)
            return U, ctx.stat(9), ctx.stat(2), ctx.stat(3), ctx.stat(10)

    U, mode, M_used, M_ref, products = run()
    assert 6 <= M_ref <= 8 and M_used == 8 and mode == 3 and products == 4
    Uo = equiprop_oracle(H0, H1, carr, dt, quad, mag, prec)
    assert rel_frobenius(U, Uo) < tol
    monkeypatch.setenv("PARAMENT_SERIES", "horner")
    V, mode_h, *_ = run()
    assert mode_h == 1
    assert rel_frobenius(U, V) < tol


@pytest.mark.parametrize("n,A,quad,mag,prec,dt,onchip", [
    (20, 2, "none", False, "fp64", 0.2, True), (32, 1, "midpoint", False, "fp64", 0.2, True),
    (40, 3, "simpson", False, "fp64", 0.1, True), (64, 2, "simpson", True, "fp64", 0.1, True),
    (64, 4, "none", False, "fp64", 0.1, True), (48, 2, "none", False, "fp32", 0.9, True),
    (40, 2, "simpson", False, "fp64", 0.1, False), (96, 2, "none", False, "fp64", 0.2, True),
    (130, 1, "simpson", False, "fp64", 0.1, True), (16, 2, "simpson", False, "fp64", 0.1, True),
    (11, 3, "simpson", True, "fp64", 0.1, True), (6, 2, "midpoint", False, "fp64", 0.2, True), (2, 1, "none", False, "fp64", 0.2, True),
    (16, 2, "none", False, "fp32", 0.9, True)])
def test_degree12_four_product_path(pb, n, A, quad, mag, prec, dt, onchip, monkeypatch):
    """Table degrees 9..12 are evaluated as one degree-12 polynomial in four matrix products in every kernel family
    (api.cu solve_degree12; k1_warp.cu, k4_onchip.cu, k4_gemm.cu build_program).  It must be active,
    meet the north_star tolerance against the oracle and agree with the Paterson-Stockmeyer / Horner evaluation.
    (PARAMENT_NORM=reference: see test_degree8_three_product_path.)"""
    monkeypatch.setenv("PARAMENT_NORM", "reference")
    rng = np.random.default_rng(1000 * n + A)
    herm = lambda: (lambda g: (g + g.conj().T) / 2)(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    norm1 = lambda m: m / np.max(np.sum(np.abs(m), axis=1))
    ctype = np.complex64 if prec == "fp32" else np.complex128
    H0 = (0.5 * norm1(herm())).astype(ctype)
    H1 = [(0.5 / A * norm1(herm())).astype(ctype) for _ in range(A)]
    pts = 41 if n > 64 else 161
    carr = (rng.uniform(-1, 1, (A, pts)) + 1j * rng.uniform(-0.3, 0.3, (A, pts))).astype(ctype)
    if not onchip:
        monkeypatch.setenv("PARAMENT_NO_ONCHIP", "1")

    def run():
        with pb.Parament(prec) as ctx:
            ctx.set_hamiltonian(H0, *H1, use_magnus=mag, quadrature_mode=quad)
            U = ctx.equiprop(dt, *carr)
            return U, ctx.stat(9), ctx.stat(2), ctx.stat(3), ctx.stat(10), ctx.stat(5)

    U, mode, M_used, M_ref, products, fam = run()
    assert fam == (1 if n <= 16 else (2 if n <= 64 else 3))
    assert 9 <= M_ref <= 12 and M_used == 12 and mode == 4 and products == 5
    Uo = equiprop_oracle(H0, H1, carr, dt, quad, mag, prec)
    assert rel_frobenius(U, Uo) < (2e-6 if prec == "fp32" else 1e-13)
    monkeypatch.setenv("PARAMENT_SERIES", "horner")
    V, mode_h, *_ = run()
    assert mode_h in (1, 2)
    assert rel_frobenius(U, V) < (2e-6 if prec == "fp32" else 1e-13)


@pytest.mark.parametrize("case", [c for c in extended_cases() + reference_test_cases()], ids=lambda c: c["name"])
def test_vs_reference_cuda_build(pb, case):
    if REF_CUDA is None or case["name"] not in REF_CUDA.files:
        pytest.skip("no reference-CUDA vector (reference undefined for this case, or ref_cuda.npz not generated)")
    U, _ = run_case(pb, case)
    N, _ = effective_steps(case["carr"].shape[1], case["quadrature"], case["use_magnus"])
    # the reference itself drifts ~1e-8 N (fp32) / 3e-17 N (fp64) from the truth
    tol = (2e-6 + 3e-8 * N) if case["precision"] == "fp32" else (1e-13 + 1e-15 * N)
    assert rel_frobenius(U, REF_CUDA[case["name"]]) < tol


@pytest.mark.parametrize("mmax", [1, 2, 3, 4, 6, 8, 11, 12, 20])
@pytest.mark.parametrize("n", [4, 16, 40])
def test_any_degree_is_a_chebyshev_truncation(pb, n, mmax):
    """Manual MMAX of either parity (the reference is only correct for odd >= 3, SURVEY A-6): the result must be
    the degree-MMAX Chebyshev/Bessel truncation  J0 I + 2 sum_k (-i)^k J_k(x) T_k(H/Hnorm)  of every step."""
    from scipy.special import jv
    rng = np.random.default_rng(n * 100 + mmax)
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H0 = (g + g.conj().T) / 2
    H0 = H0 / np.max(np.sum(np.abs(H0), axis=1))
    H1 = np.zeros((n, n))
    dt = 1.5
    with pb.Parament("fp64") as ctx:
        ctx.set_hamiltonian(H0, H1)
        ctx.set_iteration_cycles(mmax)
        U = ctx.equiprop(dt, np.zeros(3))
        assert ctx.stat(2) == mmax
        hn = ctx.stat(8)
    x = dt * hn
    Xs = H0 / hn
    T0, T1 = np.eye(n), Xs
    S = jv(0, x) * T0 + 2 * (-1j) * jv(1, x) * T1
    for k in range(2, mmax + 1):
        T0, T1 = T1, 2 * Xs @ T1 - T0
        S = S + 2 * ((-1j) ** k) * jv(k, x) * T1
    assert rel_frobenius(U, S @ S @ S) < 1e-12


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configurations
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,pts", [("C1", None), ("C2", 20001), ("C3", 1500), ("C4", 48), ("C5", None)])
def test_baseline_configs_reduced(pb, name, pts):
    w = make_workload(name, pts=pts, batch=1 if name == "C5" else None)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        U = ctx.equiprop(w.dt, *w.carr)
    Uo = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision, workers=4)
    assert rel_frobenius(U, Uo) < TOL[w.precision]


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4"])
def test_baseline_configs_full_size(pb, name):
    """Full BASELINE size against the precomputed float64 oracle propagator (tests/golden/make_full_golden.py)."""
    gold = _load(f"full_{name}.npz")
    if gold is None:
        pytest.skip(f"full_{name}.npz not generated")
    w = make_workload(name)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        U = ctx.equiprop(w.dt, *w.carr)
    assert rel_frobenius(U, gold["U"]) < TOL[w.precision]
    # size-independent property: the propagator of a Hermitian generator is unitary
    n = w.dim
    Ud = U.astype(np.complex128)
    assert np.linalg.norm(Ud.conj().T @ Ud - np.eye(n)) / np.sqrt(n) < (5e-6 if w.precision == "fp32" else 1e-11)


def test_ensemble_full_size(pb):
    """configs[4]: 1e4 independent dim-8 pulses x 1e3 points in one call."""
    w = make_workload("C5")
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1)
        U = ctx.equiprop_batch(w.dt, w.carr)
        single = ctx.equiprop(w.dt, *w.carr[1234])
    assert U.shape == (w.batch, 8, 8)
    gold = _load("full_C5.npz")
    if gold is not None:
        for b in range(gold["U"].shape[0]):
            assert rel_frobenius(U[b], gold["U"][b]) < TOL["fp32"]
    for b in (0, 777, 5000, 9999):
        Uo = equiprop_oracle(w.H0, w.H1, w.carr[b], w.dt, "none", False, "fp32")
        assert rel_frobenius(U[b], Uo) < TOL["fp32"]
    assert rel_frobenius(U[1234], single) < 5e-6      # ensemble and single call chunk the pulse differently (fp32 / TF32 arithmetic)
    Ud = U.astype(np.complex128)
    dev = np.linalg.norm(np.einsum("bji,bjk->bik", Ud.conj(), Ud) - np.eye(8), axis=(1, 2)).max()
    assert dev < 1e-5


@pytest.mark.parametrize("name,pts,slices", [("C2", 100001, 4), ("C1", 10000, 3), ("C3", 801, 2), ("C2", 30000, 8)])
def test_time_slices_compose(pb, name, pts, slices):
    """Multi-GPU decomposition: partial propagators of contiguous slices, combined in order, equal the
    whole-pulse propagator."""
    w = make_workload(name, pts=pts)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        whole = ctx.equiprop(w.dt, *w.carr)
        N = w.steps
        bounds = [N * g // slices for g in range(slices + 1)]
        parts = [ctx.equiprop_slice(w.dt, w.carr, bounds[g], bounds[g + 1]) for g in range(slices)]
        combined = ctx.combine(np.stack(parts))
    host = np.eye(w.dim, dtype=np.complex128)
    for p in parts:
        host = p.astype(np.complex128) @ host
    tol = 2e-6 if w.precision == "fp32" else 1e-12
    assert rel_frobenius(combined, whole) < tol
    assert rel_frobenius(combined, host) < tol


def test_device_resident_combine(pb):
    """Parament_combineDevice: the ordered product of slice propagators that already sit in HBM (what bench.py does with
    the NCCL all-gather output)."""
    torch = pytest.importorskip("torch")
    w = make_workload("C2", pts=40001)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="simpson")
        N = w.steps
        b = [N * g // 5 for g in range(6)]
        parts = np.stack([ctx.equiprop_slice(w.dt, w.carr, b[g], b[g + 1]) for g in range(5)])
        host = ctx.combine(parts)
        d_parts = torch.from_numpy(parts).cuda()
        d_out = torch.zeros(16, 16, dtype=torch.complex64, device="cuda")
        ctx.combine_device(d_parts.data_ptr(), 5, d_out.data_ptr())
        assert np.array_equal(d_out.cpu().numpy(), host)
        whole = ctx.equiprop(w.dt, *w.carr)
    assert rel_frobenius(host, whole) < 2e-6


def test_time_reversal_property(pb):
    """U(H, c)^-1 = U(-H, reversed c) for QUADRATURE_NONE -- a full-size, oracle-free check."""
    w = make_workload("C2", pts=200001)
    with pb.Parament("fp64") as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1)
        fwd = ctx.equiprop(w.dt, *w.carr)
        ctx.set_hamiltonian(-w.H0, *(-w.H1))
        bwd = ctx.equiprop(w.dt, *w.carr[:, ::-1])
    assert np.linalg.norm(bwd @ fwd - np.eye(16)) / 4 < 1e-11


# ---------------------------------------------------------------------------------------------------------
# edge cases and error behaviour (reference test_error.py + SURVEY App. A)
# ---------------------------------------------------------------------------------------------------------
def test_zero_effective_steps_is_identity(pb):
    with pb.Parament("fp64") as ctx:
        ctx.set_hamiltonian(np.diag([1.0, -1.0]), np.eye(2), quadrature_mode="midpoint")
        assert np.array_equal(ctx.equiprop(0.1, np.ones(1)), np.eye(2))
        ctx.set_hamiltonian(np.diag([1.0, -1.0]), np.eye(2), quadrature_mode="simpson")
        assert np.array_equal(ctx.equiprop(0.1, np.ones(2)), np.eye(2))


def test_zero_hamiltonian(pb):
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(np.zeros((3, 3)), np.zeros((3, 3)))
        assert np.array_equal(ctx.equiprop(0.1, np.ones(5)), np.eye(3))


def test_error_codes(pb):
    lib = pb._lib.lib
    with pb.Parament() as ctx:
        with pytest.raises(RuntimeError, match="No hamiltonian set"):
            ctx.amps = 1; ctx.dim = 2
            ctx.equiprop(1.0, np.zeros(3))
        assert lib.Parament_peekAtLastError(ctx._handle) == 80
        with pytest.raises(ValueError, match="Invalid quadrature selection"):
            ctx.set_hamiltonian(np.eye(2), np.eye(2), use_magnus=True, quadrature_mode="none")
        with pytest.raises(ValueError, match="Invalid quadrature selection"):
            ctx.set_hamiltonian(np.eye(2), np.eye(2), use_magnus=True, quadrature_mode="midpoint")
        assert lib.Parament_getLastError(ctx._handle) == 90
        # error 90 leaves the context without a Hamiltonian (parament.cpp:365-367)
        with pytest.raises(RuntimeError, match="No hamiltonian set"):
            ctx.equiprop(1.0, np.zeros(3))
        ctx.set_hamiltonian(np.eye(2), np.eye(2))
        with pytest.raises(RuntimeError, match="Timestep too large"):      # auto degree beyond the table: code 70
            ctx.equiprop(100.0, np.zeros(3))
        assert lib.Parament_peekAtLastError(ctx._handle) == 70
        ctx.set_iteration_cycles(7)                                          # manual degree bypasses the table
        ctx.equiprop(100.0, np.zeros(3))
        ctx.set_iteration_cycles(None)
        with pytest.raises(ValueError, match="Got 2 amplitude arrays, but there are only 1 Hamiltonians."):
            ctx.equiprop(1.0, np.zeros(4), np.zeros(4))
        with pytest.raises(ValueError, match="All amplitude arrays must have the same length."):
            ctx.set_hamiltonian(np.eye(2), np.eye(2), np.eye(2))
            ctx.equiprop(1.0, np.zeros(4), np.zeros(5))
    with pytest.raises(RuntimeError, match="Attempting to use a context that has been destroyed"):
        ctx.set_hamiltonian(np.eye(2), np.eye(2))


def test_reuse_and_reset(pb):
    """Repeated calls, growing and shrinking pulses, changing quadrature / dimension / control count on one context
    (the reference leaks or under-allocates here, SURVEY A-2, A-4)."""
    rng = np.random.default_rng(9)
    with pb.Parament("fp64") as ctx:
        for n, A, pts, quad, mag in [(4, 1, 50, "none", False), (4, 3, 500, "simpson", True), (16, 2, 20, "midpoint", False),
                                     (40, 2, 33, "simpson", False), (8, 4, 2000, "none", False), (4, 1, 5, "none", False)]:
            g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
            H0 = (g + g.conj().T) / (4 * n)
            H1 = [((h := rng.standard_normal((n, n))) + h.T) / (4 * n * A) for _ in range(A)]
            carr = rng.uniform(-1, 1, (A, pts))
            ctx.set_hamiltonian(H0, *H1, use_magnus=mag, quadrature_mode=quad)
            for _ in range(2):
                U = ctx.equiprop(0.05, *carr)
                Uo = equiprop_oracle(H0, np.array(H1), carr, 0.05, quad, mag, "fp64")
                assert rel_frobenius(U, Uo) < 1e-12


def test_magnus_many_controls(pb):
    """Magnus with 4 and 6 controls: the reference's commutator slot map collides there (SURVEY A-3); ours is the
    closed form for any number of controls."""
    rng = np.random.default_rng(11)
    for A in (4, 6):
        n = 8
        mk = lambda: ((g := rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) + g.conj().T) / (8 * n)
        H0, H1 = mk(), [mk() for _ in range(A)]
        carr = rng.uniform(-1, 1, (A, 101))
        with pb.Parament("fp64") as ctx:
            ctx.set_hamiltonian(H0, *H1, use_magnus=True, quadrature_mode="simpson")
            U = ctx.equiprop(0.1, *carr)
        assert rel_frobenius(U, equiprop_oracle(H0, np.array(H1), carr, 0.1, "simpson", True, "fp64")) < 1e-12


@pytest.mark.parametrize("n,A,pts,quad,mag", [(33, 2, 40, "none", False), (48, 3, 30, "simpson", True), (63, 1, 25, "midpoint", False),
                                              (65, 2, 9, "none", False), (100, 2, 7, "simpson", False), (17, 12, 50, "none", False),
                                              (9, 8, 61, "simpson", True), (130, 1, 5, "midpoint", False), (70, 2, 11, "simpson", True)])
def test_odd_dimensions_and_many_controls(pb, n, A, pts, quad, mag):
    """Padding of every kernel family (dims just above / below the family boundaries) and large control counts
    (Magnus with 8 controls = 44 effective terms)."""
    rng = np.random.default_rng(n * 7 + A)
    mk = lambda s: s * ((g := rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) + g.conj().T) / (2 * n)
    H0, H1 = mk(0.5), [mk(0.5 / A) for _ in range(A)]
    carr = rng.uniform(-1, 1, (A, pts))
    with pb.Parament("fp64") as ctx:
        ctx.set_hamiltonian(H0, *H1, use_magnus=mag, quadrature_mode=quad)
        U = ctx.equiprop(0.08, *carr)
    assert rel_frobenius(U, equiprop_oracle(H0, np.array(H1), carr, 0.08, quad, mag, "fp64")) < 1e-12


def test_too_many_effective_terms_is_an_error(pb):
    """Magnus with 11 controls needs 77 effective terms; the kernels take at most 64 (include/parament.h): setHamiltonian
    rejects it with code 50 (invalid value) and leaves the context without a Hamiltonian."""
    n, A = 4, 11
    rng = np.random.default_rng(0)
    H = [rng.standard_normal((n, n)) for _ in range(A + 1)]
    with pb.Parament("fp64") as ctx:
        with pytest.raises(ValueError, match="Invalid value"):
            ctx.set_hamiltonian(H[0], *H[1:], use_magnus=True, quadrature_mode="simpson")
        out = np.zeros(n * n, dtype=np.complex128)
        carr = np.ascontiguousarray(rng.uniform(-1, 1, (A, 9)).astype(np.complex128).ravel())
        assert pb._lib.lib.Parament_equiprop_fp64(ctx._handle, carr, 0.001, 9, A, out) == 80      # no Hamiltonian set


def test_device_resident_operands(pb):
    """Parament_equipropDevice: amplitudes and result stay in HBM (torch only provides the device memory)."""
    torch = pytest.importorskip("torch")
    w = make_workload("C2", pts=50001)
    carr = torch.from_numpy(w.carr).cuda()
    out = torch.zeros(16, 16, dtype=torch.complex64, device="cuda")
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="simpson")
        host = ctx.equiprop(w.dt, *w.carr)
        ctx.equiprop_device(w.dt, carr.data_ptr(), w.pts, w.amps, out.data_ptr())
        torch.cuda.synchronize()
        assert ctx.stat(1) >= 1 and ctx.stat(0) > 0       # one fused launch (chain + ordered reduction) since round 2
    assert rel_frobenius(out.cpu().numpy(), host) < 1e-7


def test_concurrent_contexts(pb):
    """Distinct contexts from distinct threads (ctypes drops the GIL; one stream per context)."""
    w = make_workload("C2", pts=20001)
    Uo = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, False, "fp32", workers=2)
    errs = []

    def work():
        with pb.Parament("fp32") as ctx:
            ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="simpson")
            for _ in range(5):
                errs.append(rel_frobenius(ctx.equiprop(w.dt, *w.carr), Uo))

    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(errs) == 20 and max(errs) < 1e-5


def test_error_growth_is_not_linear_in_steps(pb):
    """The reference's error grows ~1e-8 per step in complex64 (1e-3 at 1e5 steps).  Ours must hold
    the 1e-5 tolerance with margin as N grows (double-precision arithmetic on the tensor pipe, E-form series)."""
    errs = {}
    for pts in (2001, 20001, 200001):
        w = make_workload("C2", pts=pts)
        with pb.Parament("fp32") as ctx:
            ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="simpson")
            U = ctx.equiprop(w.dt, *w.carr)
        errs[pts] = rel_frobenius(U, equiprop_oracle(w.H0, w.H1, w.carr, w.dt, "simpson", False, "fp32", workers=4))
    assert max(errs.values()) < 2e-6, errs


@pytest.mark.parametrize("variant", [{"PARAMENT_K4_3M": "0", "PARAMENT_K4_FEED": "tma"}, {"PARAMENT_K4_3M": "0"}, {"PARAMENT_K4_3M": "1"},
                                     {"PARAMENT_K4_3M": "2"}, {"PARAMENT_K4_3M": "3", "PARAMENT_F3_STREAMS": "1"}],
                         ids=lambda v: "-".join(f"{k[9:]}={x}" for k, x in v.items()))
def test_batched_gemm_variants(tmp_path, variant):
    """The measured alternatives of the batched GEMM for dim > 64: four real products per complex product
    with the cp.async or the bulk-copy (TMA engine) + mbarrier operand feed, three real products on 64x64 / 64x32 tiles, one
    chunk stream.  The switches are read once per process, hence a subprocess; every variant must meet the tolerance."""
    import subprocess
    import sys
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {root!r})
        import parament_b200 as pb
        from workloads import make_workload
        for name, pts in (("C4", 150), ("C3", 300)):
            w = make_workload(name, pts=pts)
            H0, H1 = w.H0, w.H1
            if name == "C3":                      # C3's pulse on a dim-96 system: zero-padded to 128 in the batched pipeline
                rng = np.random.default_rng(5)
                from workloads import rand_herm
                H0 = (0.5 * rand_herm(rng, 96)).astype(np.complex128)
                H1 = np.stack([(0.125 * rand_herm(rng, 96)).astype(np.complex128) for _ in range(4)])
            with pb.Parament("fp64") as ctx:
                ctx.set_hamiltonian(H0, *H1, quadrature_mode="none")
                U = ctx.equiprop(w.dt, *w.carr)
                assert ctx.stat(5) == 3
            np.save({str(tmp_path)!r} + "/" + name + ".npy", U)
            np.savez({str(tmp_path)!r} + "/" + name + "_in.npz", H0=H0, H1=H1, carr=w.carr, dt=w.dt)
    """)
    env = dict(os.environ, **variant)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("C4", "C3"):
        U = np.load(tmp_path / f"{name}.npy")
        d = np.load(tmp_path / f"{name}_in.npz")
        Uo = equiprop_oracle(d["H0"], d["H1"], d["carr"], float(d["dt"]), "none", False, "fp64")
        assert rel_frobenius(U, Uo) < 1e-12
