"""Generate tests/golden/full_<cfg>.npz: CPU-oracle propagators of the BASELINE.json configurations at FULL size
(float64 scipy.linalg.expm product, all host cores).  Minutes of CPU time, no GPU:

    python tests/golden/make_full_golden.py C1 C2 C3 C4 C5

C5 stores the first 16 pulses of the 1e4-pulse ensemble (the generator is seeded per configuration, so the
full ensemble is reproducible and any pulse can be re-derived).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.equiprop_oracle import equiprop_oracle  # noqa: E402
from workloads import make_workload  # noqa: E402

if __name__ == "__main__":
    cores = os.cpu_count()
    for name in sys.argv[1:] or ["C1", "C2", "C5"]:
        w = make_workload(name)
        t = time.time()
        if w.batch > 1:
            U = np.stack([equiprop_oracle(w.H0, w.H1, w.carr[b], w.dt, w.quadrature, w.use_magnus, w.precision)
                          for b in range(16)])
        else:
            U = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision, workers=cores)
        dt = time.time() - t
        np.savez_compressed(os.path.join(HERE, f"full_{name}.npz"), U=U, seconds=dt, cores=cores, steps=w.steps)
        print(name, "steps", w.steps, "seconds", round(dt, 1), "cores", cores, flush=True)
