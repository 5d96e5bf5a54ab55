#!/bin/bash
# GPU session r2b: full GPU suite, TF32 error sweep, C5 arithmetic A/B
mkdir -p gpurun_out/r2b
(timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r2b/pytest.log 2>&1; tail -5 gpurun_out/r2b/pytest.log
timeout 300 python tools/gpu_errgrowth_tf32.py > gpurun_out/r2b/errgrowth_tf32.md 2> gpurun_out/r2b/errgrowth.err
cat gpurun_out/r2b/errgrowth_tf32.md; tail -c 400 gpurun_out/r2b/errgrowth.err
timeout 200 python bench.py --config C5 --configs none --steps 20 > gpurun_out/r2b/bench_C5_tf32.json 2> gpurun_out/r2b/bench_C5.err
PARAMENT_C64_MATH=f64 timeout 200 python bench.py --config C5 --configs none --steps 20 > gpurun_out/r2b/bench_C5_f64.json 2>> gpurun_out/r2b/bench_C5.err
PARAMENT_TF32_COMP=0 timeout 200 python bench.py --config C5 --configs none --steps 20 > gpurun_out/r2b/bench_C5_tf32_nocomp.json 2>> gpurun_out/r2b/bench_C5.err
tail -c 300 gpurun_out/r2b/bench_C5.err
python - <<'PY'
import json
for f in ["tf32", "f64", "tf32_nocomp"]:
    try:
        d = json.load(open("gpurun_out/r2b/bench_C5_%s.json" % f))
        print(f, "%.4g" % d["value"], d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "pinned %.4g" % d["e2e"]["pinned"]["value"], d["implementation"])
    except Exception as e:
        print(f, "failed", e)
PY
