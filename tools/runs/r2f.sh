#!/bin/bash
# GPU session r2f: on-chip kernel with three-real-product complex multiply: parity + C3 bench + ncu
O=gpurun_out/r2f; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q -x) > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python bench.py --config C3 --configs none --steps 5 > $O/bench_C3.json 2> $O/bench_C3.err; tail -c 300 $O/bench_C3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f/bench_C3.json"))
print("C3 %.4g" % d["value"], "ms %.4g" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], d["implementation"])
PY
timeout 120 python tools/gpu_fullerr.py C3 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k4_onchip -c 1 -o $O/onchip_C3 -f python tools/ncu_target.py C3 20000 1 > $O/ncu.log 2>&1; tail -2 $O/ncu.log
