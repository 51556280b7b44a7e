"""Summarise an .ncu-rep (raw page) to the handful of metrics cited in DESIGN.md / profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "sm__pipe_fmaheavy", "pipe_tensor",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
    "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "lts__t_sector_hit_rate", "smsp__warp_issue_stalled",
]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")], "grid", vals[hdr.index("Grid Size")], "block", vals[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, vals):
        if any(p in h for p in pats):
            print(f"  {h:95s} {v:>16s} {u}")
