"""Seeded synthetic Hamiltonians and control pulses for the BASELINE.json configurations.

Shapes and normalisation follow SURVEY.md section 8(d): every matrix is Hermitian and scaled so that the
reference's norm bound (`parament.cpp:280-284`, max-row-abs-sum of H0 plus that of every H_k) is
exactly 1, pulses are smooth, real and bounded by 1, and dt is chosen so that Hnorm*h = 0.2, which makes
the reference pick MMAX = 5 (complex64) / 11 (complex128) (`parament.cpp:726,750`).  numpy only: the
same generator feeds the GPU library, the CPU oracle and the reference build.  This module is bench / test
infrastructure and deliberately lives outside the parament_b200 package: importing it must not load libparament.so
(bench.py --impl reference runs the reference's library only).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Workload:
    name: str
    dim: int
    amps: int
    pts: int
    precision: str          # 'fp32' | 'fp64'
    quadrature: str         # 'none' | 'midpoint' | 'simpson'
    use_magnus: bool
    dt: float
    H0: np.ndarray
    H1: np.ndarray          # (amps, dim, dim)
    carr: np.ndarray        # (amps, pts) or (batch, amps, pts)
    batch: int = 1
    description: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def ctype(self):
        return np.complex64 if self.precision == "fp32" else np.complex128

    @property
    def steps(self) -> int:
        """effective steps per pulse (parament.cpp:820-831)"""
        if self.use_magnus or self.quadrature == "simpson":
            return max((self.pts - 1) // 2, 0)
        if self.quadrature == "midpoint":
            return max(self.pts - 1, 0)
        return self.pts

    @property
    def total_steps(self) -> int:
        return self.steps * self.batch


def _row_abs_sum(m):
    return float(np.max(np.sum(np.abs(m), axis=1)))


def rand_herm(rng, n):
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    h = (g + g.conj().T) / 2
    return h / _row_abs_sum(h)


def smooth_pulses(rng, amps, pts, batch=None, dtype=np.float64):
    """c_k(t_j) = sum_{m=1..4} a_km sin(2 pi m j/(P-1) + phi_km), sum_m a_km = 1  =>  |c_k| <= 1."""
    shape = (amps,) if batch is None else (batch, amps)
    a = rng.uniform(0.0, 1.0, shape + (4,))
    a /= a.sum(axis=-1, keepdims=True)
    phi = rng.uniform(0.0, 2 * np.pi, shape + (4,))
    j = np.arange(pts, dtype=np.float64) / max(pts - 1, 1)
    out = np.zeros(shape + (pts,), dtype=dtype)
    for m in range(4):
        out += (a[..., m, None] * np.sin(2 * np.pi * (m + 1) * j + phi[..., m, None])).astype(dtype)
    return out


_SPECS = {
    # name: (cfg index, dim, amps, pts, precision, quadrature, magnus, batch)
    "C1": (1, 2, 1, 10_000, "fp64", "midpoint", False, 1),
    "C2": (2, 16, 2, 1_000_000, "fp32", "simpson", False, 1),
    "C3": (3, 64, 4, 1_000_000, "fp64", "none", False, 1),
    "C4": (4, 256, 8, 100_000, "fp64", "none", False, 1),
    "C5": (5, 8, 2, 1_000, "fp32", "none", False, 10_000),
    # variants of C2 the north_star names but BASELINE.json has no line for: the Magnus commutator terms (A' = 5 effective
    # terms per step) and complex amplitudes (the full-cost branch of the assembly)
    "C2_magnus": (2, 16, 2, 1_000_000, "fp32", "simpson", True, 1),
    "C2_complex": (2, 16, 2, 1_000_000, "fp32", "simpson", False, 1),
}

DESCRIPTIONS = {
    "C1": "single-qubit Rabi drive dim=2, 1 control, 1e4 points, MIDPOINT, complex128",
    "C2": "4-spin NV-centre register dim=16, 2 controls, 1e6 points, complex64, SIMPSON",
    "C3": "6-qubit system dim=64, 4 controls, 1e6 points, complex128",
    "C4": "8-qubit transmon chain dim=256, 8 controls, 1e5 points, complex128",
    "C5": "GRAPE ensemble: 1e4 independent dim=8 pulses x 1e3 points, complex64",
    "C2_magnus": "C2 with use_magnus=True: dim=16, 2 controls (5 effective terms), 1e6 points, complex64, SIMPSON + Magnus commutators",
    "C2_complex": "C2 with complex-valued amplitudes: dim=16, 2 controls, 1e6 points, complex64, SIMPSON",
}


def make_workload(name: str, pts: int | None = None, batch: int | None = None, x: float = 0.2) -> Workload:
    """Build configuration `name` (C1..C5, C2_magnus, C2_complex); `pts` / `batch` override the size (reduced-N parity cases)."""
    cfg, n, A, P, prec, quad, mag, B = _SPECS[name]
    P = P if pts is None else int(pts)
    B = B if batch is None else int(batch)
    rng = np.random.default_rng(20260000 + cfg)
    if name == "C1":
        H0 = 0.5 * np.array([[1, 0], [0, -1]], dtype=np.complex128)
        H1 = 0.5 * np.array([[[0, 1], [1, 0]]], dtype=np.complex128)
        j = np.arange(P, dtype=np.float64) / max(P - 1, 1)
        carr = np.cos(2 * np.pi * 8 * j)[None, :]
    else:
        H0 = 0.5 * rand_herm(rng, n)
        H1 = np.stack([(0.5 / A) * rand_herm(rng, n) for _ in range(A)])
        carr = smooth_pulses(rng, A, P, batch=B if B > 1 else None,
                             dtype=np.float32 if prec == "fp32" else np.float64)
        if name == "C2_complex":     # |c_k| <= 1 still holds: both parts are bounded by 1 / sqrt(2)
            carr = (carr + 1j * smooth_pulses(rng, A, P, dtype=np.float32)) / np.sqrt(2.0)
    ct = np.complex64 if prec == "fp32" else np.complex128
    H0 = H0.astype(ct)
    H1 = H1.astype(ct)
    carr = carr.astype(ct)
    h_over_dt = 2.0 if (mag or quad == "simpson") else 1.0
    hn = _row_abs_sum(H0) + sum(_row_abs_sum(h) for h in H1)
    dt = x / hn / h_over_dt
    return Workload(name=name, dim=n, amps=A, pts=P, precision=prec, quadrature=quad, use_magnus=mag,
                    dt=dt, H0=H0, H1=H1, carr=carr, batch=B, description=DESCRIPTIONS[name],
                    meta={"x": x, "hnorm": hn})
