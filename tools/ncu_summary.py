"""Prints a compact per-kernel summary from an `ncu --page raw --csv` dump (used to write profiles/*.md)."""
import csv
import sys

WANT = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("occ_limit_regs", "launch__occupancy_limit_registers"),
    ("thr_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("inst_M", "smsp__inst_executed.sum"),
    ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_membar", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("stall_notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        out = []
        for label, col in WANT:
            if col in H:
                i = H.index(col)
                v = r[i]
                try:
                    f = float(v.replace(",", ""))
                    if label == "inst_M":
                        f /= 1e6
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                unit = U[i] if label.startswith("dram_") or label == "time_us" else ""
                out.append(f"{label}={v}{unit}")
        print(name + ": " + " ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
