#!/bin/bash
O=gpurun_out/r3x; mkdir -p $O
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k1_chain -c 1 -o $O/k1_C2 -f python tools/ncu_target_dev.py C2 0 1 > $O/ncu_k1.log 2>&1; tail -1 $O/ncu_k1.log
