#!/bin/bash
# check of the committed tree: full GPU suite, smoke, bench (both arms)
O=gpurun_out/r3v; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3v/bench.json"))
print("top %.4g" % d["value"], "ms %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["e2e"].get("ms_per_step"))
for k, v in d["configs"].items():
    print(k, "%.4g" % v["value"], "e2e %.4g" % v["e2e"]["value"], "frac", v.get("roofline", {}).get("frac"))
r = json.load(open("gpurun_out/r3v/bench_ref.json"))
print("ref %.4g" % r["value"], {k: "%.4g" % v["value"] for k, v in r.get("configs", {}).items() if "value" in v})
PY
