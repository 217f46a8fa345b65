"""Per-kernel table from `ncu -i X.ncu-rep --page raw --csv` output (argv[1]): time, DRAM bytes, issue rate, stalls."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
def col(n):
    return h.index(n) if n in h else None
cols = {"us": "gpu__time_duration.sum", "rdMB": "dram__bytes_read.sum", "wrMB": "dram__bytes_write.sum",
        "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "warps%": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "issue%": "smsp__issue_active.avg.pct_of_peak_sustained_active", "Minst": "smsp__inst_executed.sum",
        "thr/inst": "smsp__thread_inst_executed_per_inst_executed.ratio", "L2hit": "lts__t_sector_hit_rate.pct",
        "L1hit": "l1tex__t_sector_hit_rate.pct", "regs": "launch__registers_per_thread",
        "st_long": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "st_short": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "st_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "st_math": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "st_lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "st_nosel": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "st_branch": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "st_disp": "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "st_bar": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "fp64%": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lsu%": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"}
units = rows[1]
print("kernel".ljust(22) + " ".join(k.rjust(8) for k in cols))
for r in rows[2:]:
    name = r[col("Kernel Name")].replace("void ", "").split("(")[0][:21]
    out = []
    for k, n in cols.items():
        i = col(n)
        if i is None or not r[i]:
            out.append("-".rjust(8)); continue
        v = float(r[i].replace(",", ""))
        if k in ("rdMB", "wrMB"):
            u = units[i]
            v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        if k == "Minst": v /= 1e6
        out.append(f"{v:8.1f}")
    print(name.ljust(22) + " ".join(out))
