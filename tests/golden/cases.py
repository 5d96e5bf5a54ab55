"""Seeded parity cases shared by the golden-vector generators and the tests (numpy only)."""
from __future__ import annotations

import numpy as np


def random_hamiltonians(dim):
    """Inputs of the reference's numerics tests (test_numerics.py:24-29), same seed and draw order."""
    np.random.seed(27)
    H0 = np.random.uniform(-1, 1, (dim, dim)) / dim + 1j * np.random.uniform(-1, 1, (dim, dim)) / 2
    H0 = H0 + H0.conj().T
    H1 = np.zeros((dim, dim), dtype=complex)
    return H0, H1


def reference_test_cases():
    """The eight parametrised cases of test_numerics.py plus the docstring vector of parament.py:61-68.

    Each case: dict(name, precision, H0, H1 (A,n,n), carr (A,pts), dt, quadrature, use_magnus, kind, threshold)
    kind 'sumabs': sum|U - expected| < threshold (test_numerics.py:44-57,84-97,60-69)
    kind 'fro'   : ||U - expected||_F < threshold (test_numerics.py:100-117)
    """
    cases = []
    for prec, eps in (("fp32", np.finfo(np.float32).eps), ("fp64", np.finfo(np.float64).eps)):
        for dim in (2, 16):
            H0, H1 = random_hamiltonians(dim)
            cases.append(dict(name=f"{prec}_expm_scipy_random_{dim}", precision=prec, H0=H0, H1=H1[None],
                              carr=np.zeros((1, 1)), dt=0.01, quadrature="none", use_magnus=False,
                              kind="sumabs", threshold=float(eps * dim * dim)))
    for dim in (2, 4):
        H0, _ = random_hamiltonians(dim)   # debug_functions.expm(m): H0 = i m, H1 = m, dt = 1, one zero amplitude
        cases.append(dict(name=f"fp32_expm_debug_{dim}", precision="fp32", H0=1j * H0, H1=H0[None],
                          carr=np.zeros((1, 1)), dt=1.0, quadrature="none", use_magnus=False,
                          kind="sumabs", threshold=float(np.finfo(np.float32).eps * dim * dim)))
    for dim in (2, 16):
        H0, H1 = random_hamiltonians(dim)
        H2 = H0[:, :]
        carr1 = np.random.rand(10)     # drawn right after random_hamiltonians, as in test_numerics.py:106-107
        carr2 = np.random.rand(10)
        cases.append(dict(name=f"fp32_multi_fields_{dim}", precision="fp32", H0=H0, H1=np.stack([H1, H2]),
                          carr=np.stack([carr1, carr2]), dt=0.01, quadrature="none", use_magnus=False,
                          kind="fro", threshold=1e-6))
    cases.append(dict(name="docstring_kat", precision="fp32", H0=np.array([[1, 0], [0, -1]], dtype=complex),
                      H1=np.array([[[0, 1], [1, 0]]], dtype=complex), carr=np.zeros((1, 1)), dt=1.0,
                      quadrature="none", use_magnus=False, kind="fro", threshold=1e-6))
    return cases


def _herm(rng, n, scale):
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    h = (g + g.conj().T) / 2
    return scale * h / np.max(np.sum(np.abs(h), axis=1))


def extended_cases():
    """Cases the reference's tests do not pin: every quadrature, Magnus, ragged sizes, complex amplitudes,
    non-Hermitian generators, fewer amplitude arrays than controls, larger dims, manual degrees."""
    rng = np.random.default_rng(20260101)
    cases = []

    def add(name, precision, n, A, pts, quad, magnus, x=0.3, complex_amp=False, nonherm=False, amps_used=None, mmax=None):
        H0 = _herm(rng, n, 0.5)
        H1 = np.stack([_herm(rng, n, 0.5 / A) for _ in range(A)])
        if nonherm:
            H0 = H0 + 0.1j * _herm(rng, n, 0.5)
        Au = A if amps_used is None else amps_used
        carr = rng.uniform(-1, 1, (Au, pts))
        if complex_amp:
            carr = carr + 1j * rng.uniform(-0.3, 0.3, (Au, pts))
        hn = np.max(np.sum(np.abs(H0), axis=1)) + sum(np.max(np.sum(np.abs(h), axis=1)) for h in H1)
        hstep = 2.0 if (magnus or quad == "simpson") else 1.0
        cases.append(dict(name=name, precision=precision, H0=H0, H1=H1, carr=carr, dt=x / hn / hstep, quadrature=quad,
                          use_magnus=magnus, mmax=mmax))

    for prec in ("fp32", "fp64"):
        for quad, mag in (("none", False), ("midpoint", False), ("simpson", False), ("simpson", True)):
            tag = "magnus" if mag else quad
            add(f"{prec}_{tag}_n4_A2_p21", prec, 4, 2, 21, quad, mag)
            add(f"{prec}_{tag}_n16_A3_p202", prec, 16, 3, 202, quad, mag)
            add(f"{prec}_{tag}_n7_A1_p34", prec, 7, 1, 34, quad, mag)
        add(f"{prec}_none_n2_A1_p1", prec, 2, 1, 1, "none", False)
        add(f"{prec}_midpoint_n3_A2_p2", prec, 3, 2, 2, "midpoint", False)
        add(f"{prec}_simpson_n5_A2_p3", prec, 5, 2, 3, "simpson", False)
        add(f"{prec}_simpson_n5_A2_p4_even", prec, 5, 2, 4, "simpson", False)
        add(f"{prec}_none_n12_A2_p500_cplx", prec, 12, 2, 500, "none", False, complex_amp=True)
        add(f"{prec}_none_n8_A2_p300_nonherm", prec, 8, 2, 300, "none", False, nonherm=True)
        add(f"{prec}_none_n16_A4_p100_fewer", prec, 16, 4, 100, "none", False, amps_used=2)
        add(f"{prec}_none_n16_A2_p50_x2", prec, 16, 2, 50, "none", False, x=2.0)
        add(f"{prec}_none_n16_A2_p30_x9", prec, 16, 2, 30, "none", False, x=9.0)
        add(f"{prec}_none_n24_A2_p60", prec, 24, 2, 60, "none", False)
        add(f"{prec}_simpson_n32_A2_p41", prec, 32, 2, 41, "simpson", False)
        add(f"{prec}_none_n64_A4_p40", prec, 64, 4, 40, "none", False, x=0.2)
        add(f"{prec}_midpoint_n70_A2_p12", prec, 70, 2, 12, "midpoint", False, x=0.2)
        add(f"{prec}_none_n16_A2_p64_m9", prec, 16, 2, 64, "none", False, mmax=9)
        add(f"{prec}_none_n16_A2_p64_m13", prec, 16, 2, 64, "none", False, mmax=13)
    add("fp64_none_n128_A2_p9", "fp64", 128, 2, 9, "none", False, x=0.2)
    return cases
