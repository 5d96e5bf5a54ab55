#!/bin/bash
# GPU session r2e (2 GPUs): bench.py --gpus 2 (weak C2 top level + strong C4/C3/C2/C5), multi-device GPU tests
O=gpurun_out/r2e; mkdir -p $O
nvidia-smi -L | head -3
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 1500 $O/bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2e/bench_n2.json").read().strip().splitlines()[-1])
    def show(n, r):
        print(n, r["scaling"], "value %.4g" % r["value"], "ms %.4g" % r["ms_per_step"], "launches", r["gpu_launches"], "e2e %.4g (%.4g ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]),
              "pinned %.4g" % r["e2e"]["pinned"]["value"], "frac %.3f" % r["roofline"]["frac"])
    show("top C2", d)
    for k, v in d["configs"].items(): show(k, v)
except Exception as e:
    print("parse failed", e)
PY
(timeout 600 python -m pytest tests/test_multi_device_gpu.py tests/test_round2_gpu.py -m gpu -q) > $O/pytest_multi.log 2>&1; tail -4 $O/pytest_multi.log
