#!/bin/bash
# GPU session r3p (8 GPUs): final multi-GPU bench lines N = 8, 4, 2 (weak C2 top level + strong C4/C3/C2/C5)
O=gpurun_out/r3p; mkdir -p $O
for N in 8 4 2; do
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5) > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -c 600 $O/bench_n$N.err | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -3
done
python - <<'PY'
import json
for N in (8, 4, 2):
    try:
        d = json.loads(open("gpurun_out/r3p/bench_n%d.json" % N).read().strip().splitlines()[-1])
        def show(n, r):
            print(N, n, r["scaling"], "value %.4g" % r["value"], "ms %.4g" % r["ms_per_step"], "launches", r["gpu_launches"], "e2e %.4g (%.4g ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]),
                  "pinned %.4g" % r["e2e"]["pinned"]["value"], "frac %.3f" % r["roofline"]["frac"], r["dtype"])
        show("top C2", d)
        for k, v in d["configs"].items(): show(k, v)
    except Exception as e:
        print(N, "parse failed", e)
PY
