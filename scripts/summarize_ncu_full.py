"""Key counters of an `ncu --set full` report, one row per captured launch.
usage: python scripts/summarize_ncu_full.py report.ncu-rep > profiles/xxx.md"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size",
    "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second",
    "sm__cycles_elapsed.avg.per_second",
]


def main(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"source: `{path}` (ncu --set full --clock-control none)\n")
    print("| # | kernel | " + " | ".join(f"{k} [{units[col[k]]}]" for k in KEYS if k in col) + " |")
    print("|---|---|" + "---:|" * sum(k in col for k in KEYS))
    for n, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        print(f"| {n} | `{name}` | " + " | ".join(r[col[k]] for k in KEYS if k in col) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
