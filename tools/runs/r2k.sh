#!/bin/bash
# GPU session r2k: three-real-product complex multiply in the dim-16 chain kernel: parity + C2 A/B
O=gpurun_out/r2k; mkdir -p $O
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python bench.py --configs C2_magnus,C2_complex --steps 20 > $O/bench_3m.json 2> $O/bench.err
PARAMENT_K1_3M=0 timeout 300 python bench.py --configs none --steps 20 > $O/bench_4m.json 2>> $O/bench.err
tail -c 300 $O/bench.err
python - <<'PY'
import json
for f in ("bench_3m", "bench_4m"):
    d = json.load(open("gpurun_out/r2k/%s.json" % f))
    print(f, "value %.4g" % d["value"], "ms %.4g" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"])
    for k, v in d["configs"].items(): print("  ", k, "%.4g" % v["value"], "ms %.4g" % v["ms_per_step"])
PY
timeout 200 python tools/gpu_fullerr.py C2 2>&1 | tail -2
PARAMENT_K1_3M=0 timeout 200 python tools/gpu_fullerr.py C2 2>&1 | tail -1
