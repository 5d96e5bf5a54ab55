#!/bin/bash
O=gpurun_out/r3y; mkdir -p $O
(timeout 200 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "tf32 or C5 or ensemble or arithmetic") > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 100 python bench.py --config C5 --configs none --steps 20 --warmup 5 > $O/bench_C5.json 2> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_C5.json"))
print("C5 value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
