"""Small single-config driver for ncu captures: python tools/ncu_target.py C3 3000 [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parament_b200 as pb
from workloads import make_workload
name, pts = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kw = {"batch": int(sys.argv[4])} if len(sys.argv) > 4 else {}
w = make_workload(name, pts=pts, **kw)
with pb.Parament(w.precision) as ctx:
    ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
    for _ in range(reps):
        U = ctx.equiprop_batch(w.dt, w.carr if w.batch > 1 else w.carr[None])
    print(name, pts, ctx.stats())
