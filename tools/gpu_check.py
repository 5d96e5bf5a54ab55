"""Quick on-GPU numerics check of the library against the CPU oracle (development aid; the real parity
tests live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parament_b200 as pb
from workloads import make_workload
from oracle.equiprop_oracle import equiprop_oracle, rel_frobenius

def run(w, quad=None, magnus=None, label=""):
    quad = w.quadrature if quad is None else quad
    magnus = w.use_magnus if magnus is None else magnus
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=magnus, quadrature_mode=quad)
        t = time.time()
        U = ctx.equiprop(w.dt, *w.carr)
        wall = time.time() - t
        st = ctx.stats()
    Uo = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, quad, magnus, w.precision, workers=8)
    err = rel_frobenius(U, Uo)
    print(f"{label or w.name:28s} n={w.dim:3d} pts={w.pts:8d} {w.precision} {quad:8s} mag={int(magnus)} err={err:.3e} "
          f"dev_ms={st['device_ms']:.3f} wall_ms={wall*1e3:.2f} M={int(st['degree_used'])}/{int(st['degree_reference'])} launches={int(st['launches'])}", flush=True)
    return err

if __name__ == "__main__":
    pb.device_info()
    run(make_workload("C1"))
    for pts in (1, 2, 3, 11, 1001, 100001):
        run(make_workload("C2", pts=pts))
    run(make_workload("C2", pts=2001), quad="none")
    run(make_workload("C2", pts=2001), quad="midpoint")
    run(make_workload("C2", pts=2001), quad="simpson", magnus=True)
    w = make_workload("C2", pts=2001); w.precision = "fp64"; w.H0 = w.H0.astype(np.complex128); w.H1 = w.H1.astype(np.complex128); w.carr = w.carr.astype(np.complex128)
    run(w, label="C2-fp64")
    run(make_workload("C5", pts=1000, batch=1))
    for pts in (1, 5, 300, 3000):
        run(make_workload("C3", pts=pts))
    run(make_workload("C4", pts=40))
    w = make_workload("C5", batch=64)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1)
        Ub = ctx.equiprop_batch(w.dt, w.carr)
        print("batch stats", ctx.stats())
    errs = [rel_frobenius(Ub[b], equiprop_oracle(w.H0, w.H1, w.carr[b], w.dt, "none", False, "fp32")) for b in range(0, 64, 9)]
    print("C5 batch errs", max(errs))
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        run(make_workload("C2"))
