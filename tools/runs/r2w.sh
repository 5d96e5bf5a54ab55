#!/bin/bash
PARAMENT_LIB_DIR=$PWD/tools/runs/lib_timing timeout 120 python tools/ncu_target_dev.py C3 100000 1 2>&1 | grep -i "phase\|C3" | head -5
timeout 120 python tools/ncu_target_dev.py C3 0 2 2>&1 | grep "C3" | head -3
timeout 120 python tools/gpu_fullerr.py C3 | tail -1
