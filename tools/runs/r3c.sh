#!/bin/bash
O=gpurun_out/r3c; mkdir -p $O
timeout 300 python bench.py --config C4 --configs none --steps 3 > $O/bench_C4.json 2> $O/bench.err; tail -c 200 $O/bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3c/bench_C4.json')); print('C4 %.4g' % d['value'], 'ms %.5g' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])"
timeout 200 python tools/gpu_fullerr.py C4 | tail -1
