#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
for M in 0 1; do for NS in 1000 2000; do
PARAMENT_K1_SKEW_MODE=$M PARAMENT_K1_SKEW_NS=$NS timeout 200 python bench.py --configs none --steps 20 > $O/bench_m${M}_s$NS.json 2>> $O/bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2u/bench_m${M}_s$NS.json")); print("mode", $M, "skew ns", $NS, "device ms %.4g" % d["ms_per_step"])
PY
done; done
