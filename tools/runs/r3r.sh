#!/bin/bash
O=gpurun_out/r3r; mkdir -p $O
timeout 300 python bench.py --config C2 --configs none --steps 20 --warmup 5 > $O/bench_C2.json 2> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_C2.json"))
print("C2 value %.4g ms %.4f e2e %.4g (%.4f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
for G in 4 8; do
PARAMENT_COPY_GROUPS=$G timeout 300 python bench.py --config C2 --configs none --steps 20 --warmup 5 > $O/bench_C2_G$G.json 2>> $O/err.log
python - <<PY
import json
d = json.load(open("$O/bench_C2_G$G.json"))
print("copy groups $G: C2 e2e %.4g (%.4f ms) pinned %.4f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pinned"]["ms_per_step"]))
PY
done
timeout 300 python tools/gpu_fullerr.py C2 | tail -1
(timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "C2 or hermitian or mixed or 16 or baseline") > $O/pytest.log 2>&1; tail -2 $O/pytest.log
