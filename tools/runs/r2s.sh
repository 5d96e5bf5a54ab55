#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
for NS in 0 500 1000 2000 4000; do
PARAMENT_K1_SKEW_NS=$NS timeout 200 python bench.py --configs none --steps 20 > $O/bench_s$NS.json 2>> $O/bench.err
done
python - <<'PY'
import json
for NS in (0, 500, 1000, 2000, 4000):
    d = json.load(open("gpurun_out/r2s/bench_s%d.json" % NS))
    print("skew ns", NS, "device ms %.4g" % d["ms_per_step"])
PY
