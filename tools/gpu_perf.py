"""Device-time sweep (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parament_b200 as pb
from parament_b200.workloads import make_workload

def sweep(name, sizes, reps=3, **kw):
    for pts in sizes:
        w = make_workload(name, pts=pts, **kw)
        with pb.Parament(w.precision) as ctx:
            ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
            for r in range(reps):
                t = time.time()
                U = ctx.equiprop_batch(w.dt, w.carr if w.batch > 1 else w.carr[None])
                wall = (time.time() - t) * 1e3
                st = ctx.stats()
                steps = w.steps * w.batch
                print(f"{name} pts={pts:8d} batch={w.batch} rep={r} dev_ms={st['device_ms']:9.3f} wall_ms={wall:9.3f} steps/s(dev)={steps/st['device_ms']*1e3:.3e} "
                      f"M={int(st['degree_used'])}/{int(st['degree_reference'])} launches={int(st['launches'])}", flush=True)

if __name__ == "__main__":
    sweep("C2", [10001, 30001, 100001, 300001, 1000001])
    sweep("C5", [1000], batch=10000)
    sweep("C1", [10000])
    sweep("C3", [2000, 20000], reps=2)
    sweep("C4", [200, 1000], reps=2)
