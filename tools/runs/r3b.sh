#!/bin/bash
O=gpurun_out/r3b; mkdir -p $O
timeout 300 python bench.py --configs C5,C1,C2_magnus --steps 20 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python - <<'PY'
import json
def show(n, r):
    print(n, "value %.4g" % r["value"], "ms %.4g" % r["ms_per_step"], "e2e %.4g (%.4g ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]), "pinned %.4g (%.4g ms)" % (r["e2e"]["pinned"]["value"], r["e2e"]["pinned"]["ms_per_step"]))
d = json.load(open("gpurun_out/r3b/bench.json"))
show("top", d)
for k, v in d["configs"].items(): show(k, v)
PY
for G in 4 5 6; do PARAMENT_COPY_GROUPS=$G timeout 100 python bench.py --configs none --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('groups $G e2e ms %.4g pinned %.4g' % (d['e2e']['ms_per_step'], d['e2e']['pinned']['ms_per_step']))"; done
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_multi_device_gpu.py -m gpu -q) 2>&1 | tail -2
