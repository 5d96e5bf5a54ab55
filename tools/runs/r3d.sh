#!/bin/bash
O=gpurun_out/r3d; mkdir -p $O
timeout 300 python tools/gpu_c1_latency.py 2>&1 | tee $O/c1_latency.log | tail -8
