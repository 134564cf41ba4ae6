"""Print the key metrics of an `ncu --page raw --csv` export per kernel launch.
usage: python tools/ncu_keys.py raw.csv [extra-metric-substring ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def main(path, extra):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hi], rows[hi + 1]
    for r in rows[hi + 2:]:
        d = dict(zip(hdr, r))
        print("==", d["Kernel Name"][:110], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in hdr:
            if any(k.endswith(key) or k == key for key in KEYS) or any(e in k for e in extra):
                print(f"   {k.split('.TriageCompute.')[-1]:95s} {d[k]:>16s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
