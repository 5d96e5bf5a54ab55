"""Device-time sweep (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parament_b200 as pb
from workloads import make_workload

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

if __name__ == "__main__" and len(sys.argv) == 1:
    sweep("C2", [10001, 30001, 100001, 300001, 1000001])
    sweep("C5", [1000], batch=10000)
    sweep("C1", [10000])
    sweep("C3", [2000, 20000], reps=2)
    sweep("C4", [200, 1000], reps=2)

def e2e(name, reps=5, pinned=False, **kw):
    import torch
    w = make_workload(name, **kw)
    carr = np.ascontiguousarray(w.carr if w.batch > 1 else w.carr[None])
    if pinned:
        t = torch.from_numpy(carr).pin_memory(); carr = t.numpy()
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        fn = getattr(pb._lib.lib, "Parament_equipropBatch" + ("_fp64" if w.precision == "fp64" else ""))
        out = np.zeros((w.batch, w.dim, w.dim), dtype=w.ctype)
        best = 1e9
        for r in range(reps):
            t0 = time.time()
            ec = fn(ctx._handle, carr.reshape(-1), float(w.dt), w.pts, w.amps, w.batch, out.reshape(-1))
            best = min(best, time.time() - t0)
        print(f"e2e {name} pinned={pinned} best wall_ms={best*1e3:.3f} steps/s={w.total_steps/best:.3e} kernel-region ms={ctx.stat(0):.3f} launches={int(ctx.stat(1))}", flush=True)

if __name__ == "__main__" and len(sys.argv) == 1:
    for pin in (False, True):
        e2e("C2", pinned=pin)
        e2e("C5", pinned=pin)
        e2e("C2", pinned=pin, pts=1000000 - 1)   # even point count: strided 2-D copy path

def custom(n, A, pts, prec="fp64", reps=2):
    rng = np.random.default_rng(n)
    mk = lambda s: s * ((g := rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) + g.conj().T) / 2
    def norm1(m): return m / np.max(np.sum(np.abs(m), axis=1))
    H0 = 0.5 * norm1(mk(1.0)); H1 = [0.5 / A * norm1(mk(1.0)) for _ in range(A)]
    carr = rng.uniform(-1, 1, (A, pts))
    with pb.Parament(prec) as ctx:
        ctx.set_hamiltonian(H0, *H1)
        for r in range(reps):
            ctx.equiprop(0.2, *carr)
            st = ctx.stats()
        print(f"custom n={n} A={A} pts={pts} {prec} dev_ms={st['device_ms']:.3f} steps/s={pts/st['device_ms']*1e3:.3e} family={int(st['family'])} products={st['products_per_step']}", flush=True)

def dev(name, reps=6, **kw):
    """Device-resident timing (the library's own CUDA events around its launches), inputs rotated over > L2 of copies."""
    import torch
    w = make_workload(name, **kw)
    carr = np.ascontiguousarray(w.carr.reshape(w.batch, w.amps, w.pts))
    nbuf = min(12, max(2, int(np.ceil(160e6 / carr.nbytes)) + 1))
    ins = [torch.from_numpy(np.roll(carr, 17 * i, axis=-1).copy()).cuda() for i in range(nbuf)]
    out = torch.zeros(w.batch, w.dim, w.dim, dtype=torch.complex128 if w.precision == "fp64" else torch.complex64, device="cuda")
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        ms = []
        for r in range(reps):
            ctx.equiprop_device(w.dt, ins[r % nbuf].data_ptr(), w.pts, w.amps, out.data_ptr(), batch=w.batch)
            torch.cuda.synchronize()
            ms.append(ctx.stats()["device_ms"])
        best, med = min(ms[1:]), float(np.median(ms[1:]))
        print(f"dev {name} pts={w.pts} batch={w.batch} best_ms={best:.4f} median_ms={med:.4f} steps/s(median)={w.total_steps/med*1e3:.4e}", flush=True)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dev":
    for name in sys.argv[2:] or ["C2", "C5", "C1"]:
        dev(name)
