#!/bin/bash
O=gpurun_out/r3z; mkdir -p $O
(timeout 150 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q -x -k "C4 or 128 or 256 or 100 or 130 or 192 or 70 or 65 or batched_gemm or any_degree") > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 60 python bench.py --config C4 --configs none --steps 2 --warmup 1 > $O/bench_C4.json 2> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_C4.json"))
print("C4 value %.4g ms %.2f e2e %.4g frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
PY
