#!/bin/bash
O=gpurun_out/r3f; mkdir -p $O
(timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "packed or packing or reference_known or extended or any_degree or baseline_configs or time_slices or device_entry or combine") > $O/pytest.log 2>&1; tail -6 $O/pytest.log
echo "--- packed"; timeout 300 python tools/gpu_small_dim_rate.py 2>&1 | tee $O/rate_packed.log
echo "--- PARAMENT_K1_PACK=0"; PARAMENT_K1_PACK=0 timeout 300 python tools/gpu_small_dim_rate.py 2>&1 | tee $O/rate_unpacked.log
echo "--- latency packed"; timeout 200 python tools/gpu_c1_latency.py 2>&1 | tee $O/c1_latency.log | cut -c1-200
