"""Error of the full-size BASELINE configurations against the precomputed float64 oracle propagators (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parament_b200 as pb
from workloads import make_workload
from oracle.equiprop_oracle import rel_frobenius
for name in sys.argv[1:] or ["C1", "C2", "C3", "C4"]:
    gold = np.load(os.path.join("tests", "golden", f"full_{name}.npz"))["U"]
    w = make_workload(name)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        U = ctx.equiprop(w.dt, *w.carr)
        print(name, "series mode", int(ctx.stat(9)), "degree", int(ctx.stat(2)), "rel. Frobenius error vs oracle %.3e" % rel_frobenius(U, gold), flush=True)
