"""Condense an `ncu --page raw --csv` export into the handful of numbers we track."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]          # ncu picks the unit per column (us/ms, Mbyte/Gbyte): print it
def col(name): return hdr.index(name) if name in hdr else None
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
        ("launch__occupancy_limit_registers", "lim_regs"), ("launch__occupancy_limit_shared_mem", "lim_smem"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
kn = col("Kernel Name")
seen = set()
for r in rows[2:]:
    name = r[kn][:70]
    if name in seen: continue
    seen.add(name)
    print(name)
    out = []
    for m, short in want:
        i = col(m)
        if i is not None:
            v = r[i]
            try: v = f"{float(v.replace(',', '')):.4g}"
            except ValueError: pass
            unit = units[i] if short in ("time", "dram_rd", "dram_wr") else ""
            out.append(f"{short}={v}{unit}")
    print("   " + " ".join(out))
