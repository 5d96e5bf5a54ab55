"""Host logic of the product-saving series evaluation (parament_b200/csrc/poly_solve.hpp), on the CPU: the solved
parameters must reproduce the polynomial they were solved for, for Taylor and for Chebyshev/Bessel coefficients."""
import math
import os
import subprocess

import numpy as np
import pytest
import scipy.special

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LD = np.longdouble


@pytest.fixture(scope="module")
def solver(tmp_path_factory):
    exe = tmp_path_factory.mktemp("poly") / "poly_solve_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "parament_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "poly_solve_check.cpp"), "-o", str(exe)], check=True)

    def run(r):
        out = subprocess.run([str(exe), str(len(r) - 1)] + [np.format_float_scientific(LD(v), precision=24) for v in r],
                             capture_output=True, text=True, check=True).stdout.split()
        return None if out == ["none"] else [LD(v) for v in out]
    return run


def pmul(a, b):
    return np.convolve(np.asarray(a, dtype=LD), np.asarray(b, dtype=LD))


def padd(*ps):
    n = max(len(p) for p in ps)
    out = np.zeros(n, dtype=LD)
    for p in ps:
        out[:len(p)] += np.asarray(p, dtype=LD)
    return out


def series_coefficients(x, M):
    """r_m of E = U - I = sum_m r_m A^m, A = -i z, for the degree-M Chebyshev/Bessel truncation at x (|z| <= 1)."""
    from numpy.polynomial import chebyshev as C
    alpha = [scipy.special.jv(0, x) - 1.0] + [2.0 * scipy.special.jv(k, x) for k in range(1, M + 1)]
    c = np.zeros(M + 1, dtype=complex)
    for k in range(M + 1):
        for m, tm in enumerate(C.cheb2poly([0] * k + [1])):
            c[m] += alpha[k] * ((-1j) ** k) * tm
    return np.real(c / np.array([(-1j) ** m for m in range(M + 1)]))


def degree8_polynomial(v):
    c4, c3, d2, d1, e2, e0 = v
    y02 = pmul([0, 0, 1], [0, c3, c4])
    return padd(pmul(padd(y02, [0, d1, d2]), padd(y02, [0, 0, e2])), e0 * y02)


def degree12_polynomial(v):
    c1, c2, c3, d1, d2, d3, e2, e3, f = v
    y0 = pmul([0, 0, 0, 1], [0, c1, c2, c3])
    return padd(pmul(padd(y0, [0, d1, d2, d3]), padd(y0, [0, 0, e2, e3])), f * y0)


CASES = [("taylor", x) for x in (0.05, 0.2, 1.0)] + [("bessel", x) for x in (0.02, 0.2, 0.5, 1.0)]


@pytest.mark.parametrize("kind,x", CASES)
def test_degree8_parameters_reproduce_the_polynomial(solver, kind, x):
    r = [x ** m / math.factorial(m) for m in range(9)] if kind == "taylor" else series_coefficients(x, 8)
    v = solver(r)
    assert v is not None
    p = degree8_polynomial(v)
    for m in range(3, 9):                       # r_0 .. r_2 are added linearly by the kernel
        assert abs(p[m] - LD(r[m])) <= 1e-15 * abs(r[m]), m
    assert abs(v[5]) < 4                         # the root with the small e0


@pytest.mark.parametrize("kind,x", CASES)
def test_degree12_parameters_reproduce_the_polynomial(solver, kind, x):
    r = [x ** m / math.factorial(m) for m in range(13)] if kind == "taylor" else series_coefficients(x, 12)
    v = solver(r)
    assert v is not None
    p = degree12_polynomial(v)
    for m in range(4, 13):                      # r_3 = d1 e2 + g3 and r_0 .. r_2 are linear
        assert abs(p[m] - LD(r[m])) <= 1e-13 * abs(r[m]), m
    assert abs(v[8]) < 6                         # the solution with the smallest f (5.02 for Taylor coefficients)


def test_no_real_solution_is_reported(solver):
    assert solver([1, 1, 1, 1, 1, 1, 1, 1, -1.0]) is None          # r8 < 0
    assert solver([0.0] * 12 + [-1.0]) is None


def test_schemes_match_direct_evaluation_in_double(solver):
    """Rounding behaviour of the nested form: evaluated in float64 on a random Hermitian generator it agrees with the
    long-double evaluation of the same polynomial as well as the float64 Horner form does."""
    rng = np.random.default_rng(5)
    n, x = 12, 0.2
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H = (g + g.conj().T) / 2
    H *= 0.9 / np.max(np.sum(np.abs(H), axis=1))
    A = -1j * H
    r = series_coefficients(x, 12)
    exact = sum(np.clongdouble(r[m]) * np.linalg.matrix_power(A.astype(np.clongdouble), m) for m in range(13))
    c1, c2, c3, d1, d2, d3, e2, e3, f = [float(v) for v in solver(r)]
    g3 = float(r[3]) - d1 * e2
    I = np.eye(n)
    A2 = A @ A
    A3 = A2 @ A
    y0 = A3 @ (c3 * A3 + c2 * A2 + c1 * A)
    E = (y0 + d3 * A3 + d2 * A2 + d1 * A) @ (y0 + e3 * A3 + e2 * A2) + f * y0 + g3 * A3 + r[2] * A2 + r[1] * A + r[0] * I
    horner = np.zeros((n, n), dtype=complex)
    for m in range(12, -1, -1):
        horner = horner @ A + r[m] * I
    err_nested = np.linalg.norm((E - exact).astype(complex))
    err_horner = np.linalg.norm((horner - exact).astype(complex))
    assert err_nested < 5e-17 and err_nested < 4 * err_horner + 1e-17
