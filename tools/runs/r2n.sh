#!/bin/bash
# GPU session r2n: mixed-precision dim-16 kernel: occupancy 2 vs 3, error, parity subset
O=gpurun_out/r2n; mkdir -p $O
export PARAMENT_K1_MIXED=1
timeout 200 python bench.py --configs none --steps 20 > $O/bench_occ2.json 2> $O/bench.err
PARAMENT_K1_OCC=3 timeout 200 python bench.py --configs none --steps 20 > $O/bench_occ3.json 2>> $O/bench.err
tail -c 300 $O/bench.err
python - <<'PY'
import json
for f in ("bench_occ2", "bench_occ3"):
    d = json.load(open("gpurun_out/r2n/%s.json" % f))
    print(f, "value %.4g" % d["value"], "ms %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"])
PY
timeout 200 python tools/gpu_fullerr.py C2 2>&1 | tail -1
PARAMENT_K1_OCC=3 timeout 200 python tools/gpu_fullerr.py C2 2>&1 | tail -1
(timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_multi_device_gpu.py -m gpu -q) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
