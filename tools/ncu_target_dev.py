"""Single-config driver for ncu captures with DEVICE-RESIDENT operands (one launch per call, the launch bench.py times):
   python tools/ncu_target_dev.py C5 [pts] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import parament_b200 as pb
from workloads import make_workload
name = sys.argv[1]
pts = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
w = make_workload(name, pts=pts)
tdt = torch.complex64 if w.precision == "fp32" else torch.complex128
carr = torch.from_numpy(np.ascontiguousarray(w.carr.reshape(w.batch, w.amps, w.pts))).cuda()
out = torch.zeros(w.batch, w.dim, w.dim, dtype=tdt, device="cuda")
with pb.Parament(w.precision) as ctx:
    ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
    for _ in range(reps):
        ctx.equiprop_device(w.dt, carr.data_ptr(), w.pts, w.amps, out.data_ptr(), batch=w.batch)
    torch.cuda.synchronize()
    print(name, w.pts, ctx.stats(), "math", ctx.stat(15))
