#!/bin/bash
# GPU session r2v: final single-GPU record: full GPU suite, smoke, default bench (both arms), launch list of the bench command
O=gpurun_out/r2v; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.err
python - <<'PY'
import json
def show(n, r):
    print(n, "value %.4g" % r["value"], "ms %.4g" % r["ms_per_step"], "launches", r["gpu_launches"], "e2e %.4g (%.4g ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]),
          "pinned %.4g" % r["e2e"]["pinned"]["value"], "wrapper", r["e2e"].get("wrapper", {}).get("ms_per_pass"), "frac %.3f" % r["roofline"]["frac"], r["dtype"])
d = json.load(open("gpurun_out/r2v/bench.json"))
show("top", d)
for k, v in d["configs"].items(): show(k, v)
r = json.load(open("gpurun_out/r2v/bench_ref.json"))
print("REF", r["value"], r["e2e"].get("pinned"), r["e2e"].get("wrapper"))
for k, v in r["configs"].items(): print("  ref", k, v.get("value"), (v.get("e2e") or {}).get("wrapper"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_C2.csv python bench.py --steps 3 --warmup 3 --configs none > $O/launches_C2.out 2>&1
timeout 300 python tools/gpu_fullerr.py C1 C2 C3 C4 > $O/fullerr.log 2>&1; cat $O/fullerr.log
