#!/bin/bash
# 8 GPUs, final tree: N = 8 and N = 2 bench lines
O=gpurun_out/r3w; mkdir -p $O
for N in 8 2; do
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5) > $O/bench_n$N.json 2> $O/bench_n$N.err
done
python - <<'PY'
import json
for N in (8, 2):
    try:
        d = json.loads(open("gpurun_out/r3w/bench_n%d.json" % N).read().strip().splitlines()[-1])
        def show(n, r):
            print(N, n, r["scaling"], "value %.4g" % r["value"], "ms %.4g" % r["ms_per_step"], "e2e %.4g (%.4g ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]), "pinned %.4g" % r["e2e"]["pinned"]["value"])
        show("top C2", d)
        for k, v in d["configs"].items(): show(k, v)
    except Exception as e:
        print(N, "parse failed", e)
PY
