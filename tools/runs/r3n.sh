#!/bin/bash
O=gpurun_out/r3n; mkdir -p $O
(timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for h in 1 0; do
PARAMENT_K4_HERM=$h timeout 300 python bench.py --config C4 --configs none --steps 3 --warmup 1 > $O/bench_C4_h$h.json 2> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_C4_h$h.json"))
print("herm $h C4 value %.4g ms %.2f e2e %.4g frac %.3f launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
PY
done
timeout 300 python tools/gpu_fullerr.py C4 | tail -1
