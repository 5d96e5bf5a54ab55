#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
for OCC in 5 6 8; do
PARAMENT_TF32_OCC=$OCC timeout 200 python bench.py --config C5 --configs none --steps 20 > $O/bench_occ$OCC.json 2>> $O/bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2z/bench_occ$OCC.json")); print("tf32 occ", $OCC, "device ms %.4g" % d["ms_per_step"], "value %.4g" % d["value"])
PY
done
