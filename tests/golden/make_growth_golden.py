"""Oracle propagators of truncated BASELINE pulses, for the error-vs-N curves (tools/gpu_errgrowth.py).

    python tests/golden/make_growth_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.equiprop_oracle import equiprop_oracle  # noqa: E402
from workloads import make_workload  # noqa: E402

SIZES = {"C2": [1001, 10001, 100001, 400001], "C3": [1000, 10000, 100000, 300000], "C4": [100, 1000, 10000]}

if __name__ == "__main__":
    store = {}
    for name, sizes in SIZES.items():
        full = make_workload(name)
        for pts in sizes:
            # the FIRST pts points of the full-size pulse, so that the curve follows one physical trajectory
            carr = full.carr[:, :pts]
            U = equiprop_oracle(full.H0, full.H1, carr, full.dt, full.quadrature, full.use_magnus, full.precision, workers=os.cpu_count())
            store[f"{name}_{pts}"] = U
            print(name, pts, flush=True)
    np.savez_compressed(os.path.join(HERE, "growth.npz"), **store)
