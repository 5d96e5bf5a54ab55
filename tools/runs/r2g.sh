#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python bench.py --config C3 --configs none --steps 5 > $O/bench_C3.json 2> $O/bench_C3.err; tail -c 300 $O/bench_C3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2g/bench_C3.json"))
print("C3 %.4g" % d["value"], "ms %.4g" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], d["implementation"]["real_products_per_complex_product"])
PY
timeout 120 python tools/gpu_fullerr.py C3 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k4_onchip -c 1 -o $O/onchip_C3 -f python tools/ncu_target.py C3 20000 1 > $O/ncu.log 2>&1; tail -1 $O/ncu.log
