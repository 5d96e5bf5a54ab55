#!/bin/bash
# phase cycle counters of the mixed dim-16 kernel (library built with -DPB_PHASE_TIMING into tools/runs/lib_timing)
PARAMENT_LIB_DIR=$PWD/tools/runs/lib_timing timeout 120 python tools/ncu_target_dev.py C2 0 1 2>&1 | grep -i "phase\|C2" | head -5
PARAMENT_K1_MIXED=0 PARAMENT_LIB_DIR=$PWD/tools/runs/lib_timing timeout 120 python tools/ncu_target_dev.py C2 0 1 2>&1 | grep -i "phase\|C2" | head -5
