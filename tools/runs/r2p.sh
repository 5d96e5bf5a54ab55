#!/bin/bash
# GPU session r2p: the evidence committed under profiles/ (launch list, ncu captures of every dominant kernel, sanitizer passes,
# error tables).  Nothing timed here is a bench value.
O=gpurun_out/r2p; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_C2.csv python bench.py --steps 3 --warmup 3 --configs none > $O/launches_C2.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_C3_C5_C1.csv python bench.py --config C3 --steps 1 --warmup 3 --configs C5,C1 > $O/launches_C3.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_chain -c 1 -o $O/k1_C2 -f python tools/ncu_target_dev.py C2 0 1 > $O/ncu_k1.log 2>&1; tail -1 $O/ncu_k1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_onchip -c 1 -o $O/onchip_C3 -f python tools/ncu_target_dev.py C3 0 1 > $O/ncu_c3.log 2>&1; tail -1 $O/ncu_c3.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_tf32 -c 1 -o $O/tf32_C5 -f python tools/ncu_target_dev.py C5 0 1 > $O/ncu_c5.log 2>&1; tail -1 $O/ncu_c5.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k4_zgemm -s 40 -c 1 -o $O/zgemm_C4 -f python tools/ncu_target_dev.py C4 2000 1 > $O/ncu_c4.log 2>&1; tail -1 $O/ncu_c4.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/gpu_sanitize.py > $O/sanitize_$tool.log 2>&1; tail -3 $O/sanitize_$tool.log
done
timeout 300 python tools/gpu_errgrowth_tf32.py > $O/error_growth_tf32.md 2> $O/errgrowth_tf32.err; tail -4 $O/error_growth_tf32.md
timeout 600 python tools/gpu_errgrowth.py > $O/error_growth.md 2> $O/errgrowth.err; tail -5 $O/error_growth.md
timeout 300 python tools/gpu_fullerr.py C1 C2 C3 C4 > $O/fullerr.log 2>&1; cat $O/fullerr.log
ls -la $O
