#!/bin/bash
O=gpurun_out/r3g; mkdir -p $O
(timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
for m in 16 8 4 2; do echo "min steps per warp $m"; PARAMENT_K1_MIN_STEPS=$m timeout 300 python tools/gpu_c1_latency.py 2>&1 | tee $O/c1_latency_$m.log | tail -4 | cut -c1-200; done
