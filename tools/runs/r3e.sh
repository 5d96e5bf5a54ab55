#!/bin/bash
O=gpurun_out/r3e; mkdir -p $O
for m in 8 4 2 1; do echo "min steps per warp $m"; PARAMENT_K1_MIN_STEPS=$m timeout 300 python tools/gpu_c1_latency.py 2>&1 | tee $O/c1_latency_$m.log | tail -4 | cut -c1-230; done
