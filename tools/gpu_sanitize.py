"""Small run of every kernel family and series path for compute-sanitizer:
   compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py
   compute-sanitizer --tool racecheck python tools/gpu_sanitize.py      (shared-memory hazards: in-CTA trees, swizzled buffers)
   compute-sanitizer --tool synccheck python tools/gpu_sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parament_b200 as pb
from workloads import make_workload
from oracle.equiprop_oracle import equiprop_oracle, rel_frobenius

# C4 pts=700: the four-stream chunk pipeline with both halves of the pending list in use (cap + carry)
for name, kw in (("C2", dict(pts=4001)), ("C5", dict(pts=101, batch=40)), ("C1", dict(pts=501)), ("C3", dict(pts=41)), ("C4", dict(pts=700))):
    w = make_workload(name, **kw)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        U = ctx.equiprop_batch(w.dt, w.carr if w.batch > 1 else w.carr[None])
        c0 = w.carr[0] if w.batch > 1 else w.carr
        Uo = equiprop_oracle(w.H0, w.H1, c0, w.dt, w.quadrature, w.use_magnus, w.precision)
        print(name, "series mode", int(ctx.stat(9)), "family", int(ctx.stat(5)), "err", rel_frobenius(U[0], Uo), flush=True)

# fused final stage with several groups of CTA partials (one long dim-16 pulse, device-resident), the TF32 kernel (short dim-8
# ensemble and a single short pulse), the one-launch combine of slice partials
import ctypes
for name, kw in (("C2", dict(pts=60001)), ("C5", dict(pts=257, batch=300)), ("C5", dict(pts=900, batch=1))):
    w = make_workload(name, **kw)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        carr = w.carr if w.batch > 1 else w.carr[None]
        U = ctx.equiprop_batch(w.dt, carr)
        N = w.steps
        parts = np.stack([ctx.equiprop_slice(w.dt, carr[0], N * g // 3, N * (g + 1) // 3) for g in range(3)])
        V = ctx.combine(parts)
        Uo = equiprop_oracle(w.H0, w.H1, carr[0], w.dt, w.quadrature, w.use_magnus, w.precision)
        print(name, kw, "launches", int(ctx.stat(1)), "math", int(ctx.stat(15)), "err", rel_frobenius(U[0], Uo), "slices", rel_frobenius(V, Uo), flush=True)

# packed small systems (dim <= 4: several systems per 8 x 8 tile, unpacked by warp shuffles + tile products), single pulse and
# ensembles in both plans; non-Hermitian inputs (general layout changes) at dim 16 / 8 and dim 128 (all tiles of Y Y)
from workloads import rand_herm
rng = np.random.default_rng(4)
for dim, prec, batch, pts in ((2, "fp64", 1, 2003), (2, "fp32", 700, 37), (4, "fp64", 9, 301), (3, "fp64", 5000, 5),
                              (16, "fp32", 1, 1501), (8, "fp32", 30, 99), (128, "fp64", 1, 9)):
    ct = np.complex64 if prec == "fp32" else np.complex128
    H0 = (0.5 * rand_herm(rng, dim)).astype(ct)
    H1 = np.stack([(0.3 * rand_herm(rng, dim)).astype(ct) for _ in range(2)])
    if dim >= 8:
        G = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
        H1[1] = (H1[1] + 0.02 * G / np.linalg.norm(G, 2)).astype(ct)
    carr = rng.uniform(-1, 1, (batch, 2, pts)).astype(ct)
    with pb.Parament(prec) as ctx:
        ctx.set_hamiltonian(H0, *H1, quadrature_mode="none")
        U = ctx.equiprop_batch(0.03, carr)
        Uo = equiprop_oracle(H0, H1, carr[batch // 2], 0.03, "none", False, prec)
        print("dim", dim, prec, "pulses", batch, "pts", pts, "math", int(ctx.stat(15)), "err", rel_frobenius(U[batch // 2], Uo), flush=True)

# single-process multi-device mode on one GPU (device 0 listed three times): threads, peer copies, combine
w = make_workload("C3", pts=900)
with pb.Parament("fp64") as ctx:
    ctx.set_devices([0, 0, 0])
    ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="none")
    U = ctx.equiprop(w.dt, *w.carr)
    Uo = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, "none", False, "fp64")
    print("multi-device", int(ctx.stat(11)), "err", rel_frobenius(U, Uo), flush=True)
