#!/bin/bash
# staging-thread sweep of the pageable host path (C5: 160 MB per call, C2: 16 MB) + sanitizer on the packed kernel + new test
O=gpurun_out/r3i; mkdir -p $O
nproc
(timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "packed") > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for t in 2 4 6 8 12 16; do
  for cfg in C5 C2; do
    PARAMENT_STAGE_THREADS=$t timeout 300 python bench.py --config $cfg --configs none --steps 10 --warmup 3 > $O/bench_${cfg}_t$t.json 2> $O/err.log
    python - <<PY
import json
d = json.load(open("$O/bench_${cfg}_t$t.json"))
print("threads $t $cfg value %.4g e2e %.4g (%.3f ms) pinned %.4g" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pinned"]["value"]))
PY
  done
done
