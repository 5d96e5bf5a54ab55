#!/bin/bash
timeout 120 python tools/ncu_target_dev.py C3 0 3 2>&1 | grep "C3" | head -3
timeout 120 python tools/gpu_fullerr.py C3 | tail -1
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "degree or extended or slices or baseline or variants") 2>&1 | tail -2
