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

# single-process multi-device mode on one GPU (device 0 listed three times): threads, peer copies, combine
w = make_workload("C3", pts=900)
with pb.Parament("fp64") as ctx:
    ctx.set_devices([0, 0, 0])
    ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="none")
    U = ctx.equiprop(w.dt, *w.carr)
    Uo = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, "none", False, "fp64")
    print("multi-device", int(ctx.stat(11)), "err", rel_frobenius(U, Uo), flush=True)
