"""A/B of the copy/compute group count of the host-pointer pipeline ($PARAMENT_COPY_GROUPS)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import parament_b200 as pb
from workloads import make_workload
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
w = make_workload(name)
carr = torch.from_numpy(np.ascontiguousarray(w.carr.reshape(w.batch, w.amps, w.pts))).pin_memory().numpy()
out = np.zeros((w.batch, w.dim, w.dim), dtype=w.ctype)
for G in (1, 2, 3, 4, 6, 8):
    os.environ["PARAMENT_COPY_GROUPS"] = str(G)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        fn = ctx._fn("Parament_equipropBatch")
        ts = []
        for r in range(8):
            t = time.perf_counter()
            assert fn(ctx._handle, carr.reshape(-1), float(w.dt), w.pts, w.amps, w.batch, out.reshape(-1)) == 0
            ts.append(time.perf_counter() - t)
        print(name, "groups", G, "e2e ms median %.3f min %.3f" % (1e3 * np.median(ts[2:]), 1e3 * min(ts)), "device ms %.3f" % ctx.stat(0), flush=True)
# raw H2D bandwidth of this box for reference
d = torch.empty(carr.nbytes, dtype=torch.uint8, device="cuda")
src = torch.from_numpy(carr.view(np.uint8).reshape(-1))
torch.cuda.synchronize()
for _ in range(3):
    t = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("pinned H2D %.1f MB in %.3f ms = %.1f GB/s" % (carr.nbytes / 1e6, dt * 1e3, carr.nbytes / dt / 1e9))
