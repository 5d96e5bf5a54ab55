"""Oracle propagators of single dim-8 complex64 pulses of growing length (the C5 Hamiltonians, one pulse of P points), for
the error-vs-N curve of the FP32 / 3xTF32 kernel against the FP64 kernel (tools/gpu_errgrowth_tf32.py).

    python tests/golden/make_growth_dim8_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.equiprop_oracle import equiprop_oracle  # noqa: E402
from workloads import make_workload  # noqa: E402

SIZES = [100, 300, 1000, 3000, 10000, 30000, 100000, 300000, 1000000]

if __name__ == "__main__":
    store = {}
    for pts in SIZES:
        w = make_workload("C5", pts=pts, batch=1)
        store[f"P{pts}"] = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision, workers=os.cpu_count())
        print(pts, flush=True)
    np.savez_compressed(os.path.join(HERE, "growth_dim8.npz"), **store)
