#!/bin/bash
# GPU session r2o: mixed precision as default: full GPU suite + default bench + reference arm
O=gpurun_out/r2o; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.err
python - <<'PY'
import json
def show(n, r):
    print(n, "value %.4g" % r["value"], "ms %.4g" % r["ms_per_step"], "launches", r["gpu_launches"], "e2e %.4g (%.4g ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]),
          "pinned %.4g" % r["e2e"]["pinned"]["value"], "wrapper", r["e2e"].get("wrapper", {}).get("ms_per_pass"), "frac %.3f" % r["roofline"]["frac"], r["dtype"])
d = json.load(open("gpurun_out/r2o/bench.json"))
show("top", d)
for k, v in d["configs"].items(): show(k, v)
r = json.load(open("gpurun_out/r2o/bench_ref.json"))
print("REF", r["value"], r["e2e"].get("pinned"), r["e2e"].get("wrapper"))
PY
timeout 300 python tools/gpu_errgrowth.py > $O/error_growth.md 2> $O/errgrowth.err; head -8 $O/error_growth.md
