#!/bin/bash
# GPU session r2m: mixed-precision dim-16 chain kernel (FP64 + TF32): error vs N and C2 bench A/B
O=gpurun_out/r2m; mkdir -p $O
export PARAMENT_K1_MIXED=1
timeout 200 python tools/gpu_fullerr.py C2 2>&1 | tail -1
timeout 300 python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import parament_b200 as pb
from workloads import make_workload
g = np.load("tests/golden/growth.npz")
full = make_workload("C2")
for pts in (1001, 10001, 100001, 400001):
    carr = np.ascontiguousarray(full.carr[:, :pts])
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(full.H0, *full.H1, quadrature_mode="simpson")
        U = ctx.equiprop(full.dt, *carr)
    G = g["C2_%d" % pts]
    print("mixed C2 pts", pts, "err %.3e" % (np.linalg.norm(U.astype(np.complex128) - G) / np.linalg.norm(G)), flush=True)
PY
timeout 300 python bench.py --configs C2_magnus,C2_complex --steps 20 > $O/bench_mixed.json 2> $O/bench.err
tail -c 300 $O/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m/bench_mixed.json"))
print("mixed", "value %.4g" % d["value"], "ms %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"])
for k, v in d["configs"].items(): print("  ", k, "%.4g" % v["value"], "ms %.4g" % v["ms_per_step"])
PY
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "C2 or extended or slices or ensemble or baseline or reference or growth") > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_chain -c 1 -o $O/k1_mixed -f python tools/ncu_target_dev.py C2 0 1 > $O/ncu.log 2>&1; tail -1 $O/ncu.log
