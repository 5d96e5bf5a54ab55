#!/bin/bash
# GPU session r2r: copy-group count of the host-pointer pipeline for C2 (e2e), smoke(), final default bench for the record
O=gpurun_out/r2r; mkdir -p $O
for G in 2 3 4 6 8; do
PARAMENT_COPY_GROUPS=$G timeout 200 python bench.py --configs none --steps 20 > $O/bench_g$G.json 2>> $O/bench.err
done
python - <<'PY'
import json
for G in (2, 3, 4, 6, 8):
    d = json.load(open("gpurun_out/r2r/bench_g%d.json" % G))
    print("groups", G, "device ms %.4g" % d["ms_per_step"], "e2e ms %.4g" % d["e2e"]["ms_per_step"], "pinned ms %.4g" % d["e2e"]["pinned"]["ms_per_step"])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
