"""Condenses an .ncu-rep (ncu --set full) into a small CSV of the metrics DESIGN.md / bench.py quote:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_name.csv
One row per captured launch.  Needs `ncu` on PATH (no GPU needed to read a report)."""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_global_red.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg",
    "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(raw.splitlines()))
    H, units, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    cols = [m for m in METRICS if m in H]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[H.index(m)]}]" for m in cols])
        for r in data:
            w.writerow([r[ki].split("(")[0]] + [r[H.index(m)] for m in cols])
    print(f"{out}: {len(data)} launches, {len(cols)} metrics")


if __name__ == "__main__":
    main()
