"""Condenses `ncu --page raw --csv` output (one row per captured launch) to the metrics DESIGN.md
and bench.py quote, and writes profiles/traffic.json (DRAM bytes per launch, per kernel):
    python tools/ncu_raw_summary.py gpurun_out/prof_TAG_raw.csv profiles/rN_name.csv [profiles/traffic.json]"""
import csv
import json
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "sm__instruction_throughput.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
    "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_global_red.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    H, U, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    cols = [m for m in METRICS if m in H]
    traffic = {}
    issue = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{U[H.index(m)]}]" for m in cols])
        for r in data:
            name = r[ki].split("(")[0].replace("void ", "").replace("gft::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
            w.writerow([name] + [r[H.index(m)] for m in cols])
            try:
                b = sum(float(r[H.index(m)].replace(",", "")) * UNIT[U[H.index(m)]]
                        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                traffic.setdefault(name.split("<")[0].replace("_kernel", "").replace("_warp", ""), []).append(b)
            except Exception:
                pass
            try:
                key = name.split("<")[0].replace("_kernel", "").replace("_warp", "")
                issue.setdefault(key, []).append(float(r[H.index("smsp__issue_active.avg.per_cycle_active")]))
            except Exception:
                pass
    print(f"{out}: {len(data)} launches, {len(cols)} metrics")
    if len(sys.argv) > 3:
        doc = {k: int(sum(v) / len(v)) for k, v in traffic.items()}
        doc["_issue_slots_busy_per_active_cycle"] = {k: round(sum(v) / len(v), 3) for k, v in issue.items()}
        doc["_source"] = src
        json.dump(doc, open(sys.argv[3], "w"), indent=1)
        print(sys.argv[3], {k: int(sum(v) / len(v)) for k, v in traffic.items()})


if __name__ == "__main__":
    main()
