"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print("=" * 100)
    for k in keys:
        if k in d: print(f"{k:75s} {u[k]:16s} {d[k]}")
    stalls = sorted(((float(v), h) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v), reverse=True)
    for v, h in stalls[:8]:
        print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
