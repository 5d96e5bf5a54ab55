#!/bin/bash
O=gpurun_out/r3l; mkdir -p $O
for h in 1 0; do
for cfg in C2; do
PARAMENT_K1_HERM=$h timeout 300 python bench.py --config $cfg --configs none --steps 20 --warmup 5 > $O/bench_${cfg}_h$h.json 2> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_${cfg}_h$h.json"))
print("herm $h $cfg value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
done; done
PARAMENT_K1_MIXED=0 PARAMENT_K1_HERM=1 timeout 300 python bench.py --config C2 --configs none --steps 20 --warmup 5 > $O/bench_C2_f64_h1.json 2>> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_C2_f64_h1.json"))
print("all-FP64 herm 1 value %.4g ms %.4f" % (d["value"], d["ms_per_step"]))
PY
(timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python tools/gpu_fullerr.py C2 | tail -2
