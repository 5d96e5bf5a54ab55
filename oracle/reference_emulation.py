"""numpy emulation of the reference's *algorithm* (not just its maths) -- TEST INFRASTRUCTURE ONLY.

Restates, step for step and in the context precision, what `/root/reference/src/cuda/parament.cpp`
does on the GPU, so that tests can ask "would the reference itself have produced this?" without a
GPU, and so that the error-vs-N behaviour of the reference's Chebyshev recurrence can be compared
with the new kernels' ("Numerics").  It is not bit-exact with cuBLAS (summation order inside
a GEMM differs); it reproduces the reference to rounding.

  * iteration-count tables ............ parament.cpp:723-766 (API-visible data, same thresholds)
  * Bessel coefficients (-i)^k J_k(x) . mathhelper.cpp:35-80, parament.cpp:404-407
  * Hnorm, alpha/beta ................. parament.cpp:280-287
  * quadrature coefficient arrays ..... control_expansion.cu:27-60,105-160
  * Magnus commutator slots ........... parament.cpp:289-359 (reproduced only for <= 3 controls,
                                        where the reference's slot map is injective)
  * assembly X_j ...................... parament.cpp:491-554
  * Clenshaw recurrence ............... parament.cpp:569-652
  * pairwise ordered tree product ..... parament.cpp:657-718

The reference works on the transposes of the row-major buffers the wrapper passes (column-major
cuBLAS); polynomials commute with transposition, so this emulation runs in the physical (row-major)
picture and multiplies `later @ earlier`.
"""
from __future__ import annotations

import numpy as np
from scipy.special import jv

from .equiprop_oracle import (QUADRATURE_MIDPOINT, QUADRATURE_NONE, QUADRATURE_SIMPSON, _QUAD_NAMES,
                              effective_dt, effective_steps, hnorm)

_FP32_THRESHOLDS = [
    (0.032516793, 3), (0.219062571, 5), (0.619625593, 7), (1.218059203, 9), (1.979888284, 11),
    (2.873301187, 13), (3.872963682, 15), (4.959398466, 17), (6.117657121, 19), (7.336154907, 21),
    (8.605792444, 23), (9.919320831, 25), (11.27088616, 27), (12.65570085, 29),
]
_FP64_THRESHOLDS = [
    (0.000213616, 3), (0.00768149, 5), (0.0501474, 7), (0.162592, 9), (0.368382, 11), (0.676861, 13),
    (1.08784, 15), (1.59605, 17), (2.19402, 19), (2.87366, 21), (3.62716, 23), (4.44725, 25),
    (5.3274, 27), (6.26179, 29), (7.2453, 31), (8.27338, 33), (9.34206, 35), (10.4478, 37),
    (11.5875, 39), (12.7584, 41),
]


def select_iteration_cycles(Hnorm: float, dt: float, precision: str) -> int:
    """parament.cpp:723-766: first table entry with Hnorm*dt <= threshold, else -1."""
    table = _FP32_THRESHOLDS if precision == "fp32" else _FP64_THRESHOLDS
    x = Hnorm * dt
    for thr, m in table:
        if x <= thr:
            return m
    return -1


def bessel_coefficients(x: float, mmax: int, ctype) -> np.ndarray:
    """J[k] = (-i)^k J_k(x), k = 0..mmax, rounded to the context type (mathhelper.cpp:69-80)."""
    k = np.arange(mmax + 1)
    return (((-1j) ** k) * jv(k, x)).astype(ctype)


def reference_equiprop_emulated(H0, H1, carr, dt, quadrature="none", use_magnus=False,
                                precision="fp32", mmax=None):
    q = _QUAD_NAMES[quadrature]
    ct = np.complex64 if precision == "fp32" else np.complex128
    rt = np.float32 if precision == "fp32" else np.float64
    H0 = np.asarray(H0).astype(ct)
    H1 = np.asarray(H1)
    if H1.ndim == 2:
        H1 = H1[None]
    H1 = H1.astype(ct)
    carr = np.atleast_2d(np.asarray(carr)).astype(ct)
    n = H0.shape[0]
    A_set = H1.shape[0]
    A = carr.shape[0]
    pts = carr.shape[1]
    Hn = hnorm(H0, H1)
    h = effective_dt(dt, q, use_magnus)
    if mmax is None:
        mmax = select_iteration_cycles(Hn, h, precision)
        if mmax < 3:
            raise RuntimeError("Timestep too large")  # parament.cpp:386-388
    J = bessel_coefficients(h * Hn, mmax, ct)
    N, _ = effective_steps(pts, q, use_magnus)
    idx = np.arange(N)

    # coefficient expansion in the context precision
    if use_magnus or q == QUADRATURE_SIMPSON:
        c0, c1, c2 = carr[:, 2 * idx], carr[:, 2 * idx + 1], carr[:, 2 * idx + 2]
        chat = ((c0 + ct(4) * c1) + c2) / ct(6)
    elif q == QUADRATURE_MIDPOINT:
        chat = ct(0.5) * (carr[:, idx] + carr[:, idx + 1])
    else:
        chat = carr[:, idx]
    mats = [H1[a] for a in range(A)]
    coefs = [chat[a] for a in range(A)]
    if use_magnus:
        if A_set > 3 or A != A_set:
            raise NotImplementedError("reference Magnus slot map is only well defined for <= 3 controls")
        fac = ct(1j * rt(h) / rt(12.0))
        for a in range(A):
            mats.append((H0 @ H1[a] - H1[a] @ H0).astype(ct))
            coefs.append((c2[a] - c0[a]) * fac)
        for k in range(A):
            for j in range(k):
                mats.append((H1[j] @ H1[k] - H1[k] @ H1[j]).astype(ct))
                coefs.append((c0[j] * c2[k] - c2[j] * c0[k]) * fac)
    X = np.broadcast_to(H0, (N, n, n)).astype(ct)
    if mats:
        X = X + np.einsum("aj,aik->jik", np.asarray(coefs, dtype=ct), np.asarray(mats, dtype=ct)).astype(ct)

    # Clenshaw, two k per trip, exactly as parament.cpp:588-648 (odd mmax only)
    if mmax % 2 == 0 or mmax < 3:
        raise NotImplementedError("the reference recurrence is only defined for odd MMAX >= 3")
    sigma = ct(2.0 / (2.0 * Hn) * 2.0)
    eye = np.eye(n, dtype=ct)
    D0 = np.zeros((N, n, n), dtype=ct)
    D1 = np.zeros((N, n, n), dtype=ct)
    k = mmax
    acc = ct(0)
    while k >= 0:
        if k == mmax:
            D0 = np.zeros_like(D0)
        else:
            D0 = (sigma * (X @ D1) - D0).astype(ct)
        D0 = D0 + J[k] * eye
        k -= 1
        if k == mmax - 1:
            acc = ct(0)
        if k == 0:
            acc = ct(-2)
        D1 = (sigma * (X @ D0) + acc * D1).astype(ct)
        D1 = D1 + J[k] * eye
        if k == mmax - 1:
            acc = ct(-1)
        k -= 1
    U = D1
    if N == 0:
        return np.eye(n, dtype=ct)
    while U.shape[0] > 1:
        m = U.shape[0] // 2
        prod = (U[1:2 * m:2] @ U[0:2 * m:2]).astype(ct)
        if U.shape[0] % 2:
            prod = np.concatenate([prod, U[-1:]], axis=0)
        U = prod
    return U[0]
