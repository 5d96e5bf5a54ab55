#!/bin/bash
# evidence for profiles/ from the final tree: sanitizer passes, ncu captures of the changed kernels, launch lists.  Nothing timed here is a bench value.
O=gpurun_out/r3q; mkdir -p $O
(timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "tf32 or hermitian or packed or spectral") > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/gpu_sanitize.py > $O/sanitize_$tool.log 2>&1; tail -2 $O/sanitize_$tool.log
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_chain -c 1 -o $O/k1_C2 -f python tools/ncu_target_dev.py C2 0 1 > $O/ncu_k1.log 2>&1; tail -1 $O/ncu_k1.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_tf32 -c 1 -o $O/tf32_C5 -f python tools/ncu_target_dev.py C5 0 1 > $O/ncu_c5.log 2>&1; tail -1 $O/ncu_c5.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k4_zgemm -s 40 -c 1 -o $O/zgemm_C4 -f python tools/ncu_target_dev.py C4 2000 1 > $O/ncu_c4.log 2>&1; tail -1 $O/ncu_c4.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_chain -c 1 -o $O/k1_C1 -f python tools/ncu_target_dev.py C1 0 1 > $O/ncu_c1.log 2>&1; tail -1 $O/ncu_c1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_C2.csv python bench.py --steps 3 --warmup 3 --configs none > $O/launches_C2.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_C4.csv python bench.py --config C4 --steps 1 --warmup 1 --configs none > $O/launches_C4.out 2>&1
ls -la $O
