import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import parament_b200 as pb
from workloads import make_workload
for name in sys.argv[1:]:
    w = make_workload(name)
    carr = torch.from_numpy(np.ascontiguousarray(w.carr)).cuda()
    out = torch.zeros(w.batch, w.dim, w.dim, dtype=torch.complex64 if w.precision == "fp32" else torch.complex128, device="cuda")
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        ms = []
        for _ in range(6):
            ctx.equiprop_device(w.dt, carr.data_ptr(), w.pts, w.amps, out.data_ptr(), batch=w.batch)
            ms.append(ctx.stat(0))
        print(f"rounds={os.environ.get('PARAMENT_K1_WAVES','1')} {name} device ms min {min(ms):.4f} median {sorted(ms)[3]:.4f} launches {int(ctx.stat(1))}", flush=True)
