"""Device-resident step rate of small systems (dim 2, 3, 4, 8) on the GPU box, single long pulse and ensemble, complex128 and complex64.
A/B of the packed kernel: run once with PARAMENT_K1_PACK=0 and once without."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import parament_b200 as pb  # noqa: E402
from workloads import rand_herm  # noqa: E402


def rate(dim, prec, batch, pts, reps=5):
    rng = np.random.default_rng(dim)
    ct = np.complex64 if prec == "fp32" else np.complex128
    H0 = (0.5 * rand_herm(rng, dim)).astype(ct)
    H1 = np.stack([(0.25 * rand_herm(rng, dim)).astype(ct) for _ in range(2)])
    carr = torch.from_numpy(rng.uniform(-1, 1, (batch, 2, pts)).astype(ct)).cuda()
    out = torch.zeros(batch, dim, dim, dtype=torch.complex64 if prec == "fp32" else torch.complex128, device="cuda")
    with pb.Parament(prec) as ctx:
        ctx.set_hamiltonian(H0, *H1, quadrature_mode="none")
        best = 1e30
        for r in range(reps + 2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.equiprop_device(0.01, carr.data_ptr(), pts, 2, out.data_ptr(), batch=batch)
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                best = min(best, e0.elapsed_time(e1))
        math = ctx.stat(15)
        deg = ctx.stat(2)
    return batch * pts / (best * 1e-3), best, math, deg


if __name__ == "__main__":
    for dim in (2, 3, 4, 8):
        for prec in ("fp64", "fp32"):
            for batch, pts in ((1, 2_000_000), (20_000, 1000)):
                r, ms, math, deg = rate(dim, prec, batch, pts)
                print(f"dim {dim} {prec} batch {batch:6d} pts {pts:8d}: {r:10.4g} steps/s  {ms:8.3f} ms  math {math:.0f} degree {deg:.0f}", flush=True)
