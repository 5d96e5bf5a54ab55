#!/bin/bash
O=gpurun_out/r3m; mkdir -p $O
for h in 1 0; do
PARAMENT_K1_HERM=$h timeout 300 python bench.py --config C5 --configs none --steps 20 --warmup 5 > $O/bench_C5_h$h.json 2> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_C5_h$h.json"))
print("herm $h C5 value %.4g ms %.4f e2e %.4g frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
PY
done
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python tools/gpu_errgrowth_tf32.py 2>&1 | tail -25
