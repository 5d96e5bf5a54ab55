#!/bin/bash
O=gpurun_out/r3u; mkdir -p $O
for cfg in C2 C2_magnus C2_complex; do
timeout 300 python bench.py --config $cfg --configs none --steps 20 --warmup 5 > $O/bench_$cfg.json 2>> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_$cfg.json"))
print("$cfg value %.4g ms %.4f e2e %.4g (%.4f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
done
(timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -2 $O/pytest.log
