#!/bin/bash
# GPU session r2c: full GPU suite, TF32 error sweep, C5 arithmetic A/B, ncu of the TF32 kernel
O=gpurun_out/r2c; mkdir -p $O
(timeout 600 python -m pytest tests -m gpu -q) > $O/pytest.log 2>&1; tail -8 $O/pytest.log
timeout 300 python tools/gpu_errgrowth_tf32.py > $O/errgrowth_tf32.md 2> $O/errgrowth.err
cat $O/errgrowth_tf32.md; tail -c 400 $O/errgrowth.err
timeout 200 python bench.py --config C5 --configs none --steps 20 > $O/bench_C5_tf32.json 2> $O/bench_C5.err
PARAMENT_C64_MATH=f64 timeout 200 python bench.py --config C5 --configs none --steps 20 > $O/bench_C5_f64.json 2>> $O/bench_C5.err
PARAMENT_TF32_COMP=0 timeout 200 python bench.py --config C5 --configs none --steps 20 > $O/bench_C5_tf32_nocomp.json 2>> $O/bench_C5.err
tail -c 300 $O/bench_C5.err
python - <<'PY'
import json
for f in ["tf32", "f64", "tf32_nocomp"]:
    try:
        d = json.load(open("gpurun_out/r2c/bench_C5_%s.json" % f))
        print(f, "%.4g" % d["value"], d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "pinned %.4g" % d["e2e"]["pinned"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_tf32 -c 1 -o $O/tf32_C5 -f python tools/ncu_target.py C5 1000 1 > $O/ncu.log 2>&1; tail -2 $O/ncu.log
