#!/bin/bash
# final check of the committed tree: full GPU suite, smoke, short bench
O=gpurun_out/r3a; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3a/bench.json"))
print("top %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], {k: "%.4g" % v["value"] for k, v in d["configs"].items()})
PY
