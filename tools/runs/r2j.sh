#!/bin/bash
# GPU session r2j: staging threads v2 (chunk queue, streaming stores): full suite + default bench + A/B without staging
O=gpurun_out/r2j; mkdir -p $O
nproc
(timeout 900 python -m pytest tests -m gpu -q) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err

PARAMENT_STAGE_THREADS=0 timeout 300 python bench.py --configs C5 --steps 10 > $O/bench_nostage.json 2> $O/bench_nostage.err
python - <<'PY'
import json
def show(n, r):
    print(n, "value %.4g" % r["value"], "ms %.4g" % r["ms_per_step"], "launches", r["gpu_launches"], "e2e %.4g (%.4g ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]),
          "pinned %.4g" % r["e2e"]["pinned"]["value"], "wrapper", r["e2e"].get("wrapper", {}).get("ms_per_pass"), "frac %.3f" % r["roofline"]["frac"])
for f in ("bench", "bench_nostage"):
    d = json.load(open("gpurun_out/r2j/%s.json" % f))
    show(f + " top", d)
    for k, v in d["configs"].items(): show(k, v)
PY
