#!/bin/bash
# GPU session r2d: full GPU suite (fused single-launch final stage, packaging), TF32 sweep, C5 / C1 / C2 bench
O=gpurun_out/r2d; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q) > $O/pytest.log 2>&1; tail -6 $O/pytest.log
timeout 300 python tools/gpu_errgrowth_tf32.py > $O/errgrowth_tf32.md 2> $O/errgrowth.err
cat $O/errgrowth_tf32.md; tail -c 400 $O/errgrowth.err
timeout 300 python bench.py --config C5 --configs C1,C2 --steps 20 > $O/bench_C5.json 2> $O/bench_C5.err
PARAMENT_C64_MATH=f64 timeout 200 python bench.py --config C5 --configs none --steps 20 > $O/bench_C5_f64.json 2>> $O/bench_C5.err
tail -c 300 $O/bench_C5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2d/bench_C5.json"))
def show(n, r):
    print(n, "%.4g" % r["value"], "ms %.4g" % r["ms_per_step"], "launches", r["gpu_launches"], "e2e %.4g (%.4g ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]),
          "pinned %.4g" % r["e2e"]["pinned"]["value"], "wrapper", r["e2e"].get("wrapper", {}).get("ms_per_pass"))
show("C5 tf32", d)
for k, v in d["configs"].items(): show(k, v)
show("C5 f64", json.load(open("gpurun_out/r2d/bench_C5_f64.json")))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_tf32 -c 1 -o $O/tf32_C5 -f python tools/ncu_target.py C5 1000 1 > $O/ncu.log 2>&1; tail -2 $O/ncu.log
