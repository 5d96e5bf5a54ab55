#!/bin/bash
O=gpurun_out/r3s; mkdir -p $O
for G in 6 5 4 3 2 4 6; do
PARAMENT_COPY_GROUPS=$G timeout 300 python bench.py --config C2 --configs none --steps 30 --warmup 5 > $O/bench_C2_G$G.json 2>> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_C2_G$G.json"))
print("copy groups $G: C2 e2e %.4g (%.4f ms) pinned %.4f ms  device %.4f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pinned"]["ms_per_step"], d["ms_per_step"]))
PY
done
