"""Summarise an .ncu-rep: `python tools/ncu_summary.py report.ncu-rep [substring ...]` (reads it with ncu -i)."""
import csv, subprocess, sys
rep = sys.argv[1]
keys = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size", "launch__waves",
    "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct", "l1tex__data_bank_conflicts",
    "smsp__warps_eligible.avg.per_cycle", "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed", "shared_mem", "sm__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in keys):
            print(f"  {h} [{u}] = {v}")
