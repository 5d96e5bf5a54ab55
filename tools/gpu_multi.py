"""End-to-end timing of the single-process multi-GPU mode (Parament_setDevices): host arrays in, propagator out, through
the host-pointer C-ABI call the unchanged wrapper makes.  STRONG scaling: the same pulse on 1, 2, ... devices.

    python tools/gpu_multi.py [C2 C3 C4 C5] [--reps 5] [--out gpurun_out/multi.json]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C2", "C3", "C5"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--pts", type=int, default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import parament_b200 as pb
    from parament_b200 import constants as K
    from workloads import make_workload
    ndev = torch.cuda.device_count()
    counts = [g for g in (1, 2, 4, 8) if g <= ndev]
    rows = []
    for name in args.configs:
        w = make_workload(name, pts=args.pts if name != "C5" else None)
        carr = np.ascontiguousarray(w.carr.reshape(w.batch, w.amps, w.pts))
        pinned = torch.from_numpy(carr).pin_memory().numpy()
        out = np.zeros((w.batch, w.dim, w.dim), dtype=w.ctype)
        base = None
        for g in counts:
            with pb.Parament(w.precision) as ctx:
                ctx.set_devices(g)
                ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
                fn = ctx._fn("Parament_equipropBatch")

                def call():
                    assert fn(ctx._handle, pinned.reshape(-1), float(w.dt), w.pts, w.amps, w.batch, out.reshape(-1)) == 0

                for _ in range(2):
                    call()
                ts = []
                for _ in range(args.reps):
                    t = time.perf_counter()
                    call()
                    ts.append(time.perf_counter() - t)
                used = int(ctx.stat(K.STAT_DEVICES_USED))
            ms = 1e3 * float(np.median(ts))
            if g == 1:
                base = out.copy()
            err = float(np.linalg.norm(out - base) / np.linalg.norm(base))
            rows.append({"config": name, "devices": g, "devices_used": used, "ms_per_call": ms,
                         "steps_per_s": w.total_steps / (ms * 1e-3), "rel_diff_vs_one_device": err})
            print(json.dumps(rows[-1]), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
