#!/usr/bin/env python3
"""Turns an ncu report (.ncu-rep, captured with --set full) into the markdown summary kept under profiles/.

usage: tools/ncu_summary.py REPORT.ncu-rep "title" "command line that produced it" > profiles/xyz.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sm__cycles_elapsed.avg",
    "sm__cycles_active.avg",
    "sm__cycles_active.min",
    "sm__cycles_active.max",
]


def main():
    rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    name = launches[0][col["Kernel Name"]]
    print(f"# {title}\n")
    print(f"Kernel: `{name}`\n")
    print(f"Command:\n\n    {cmd}\n")
    print("| metric | unit | " + " | ".join(f"launch {k + 1}" for k in range(len(launches))) + " |")
    print("|---|---|" + "---|" * len(launches))
    for m in METRICS:
        if m in col:
            print(f"| `{m}` | {units[col[m]]} | " + " | ".join(r[col[m]] for r in launches) + " |")
    # executed SASS opcode histogram of the first launch
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    hist, total, seen, h = {}, 0, 0, None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "Kernel Name":
            seen += 1
            continue
        if r and r[0] == "Address":
            h = {n: i for i, n in enumerate(r)}
            continue
        if seen != 1 or h is None or len(r) < 6:
            continue
        ins = r[1].split()
        op = (ins[1] if ins[0].startswith("@") else ins[0]).split(".")[0]
        n = int(r[h["Instructions Executed"]])
        hist[op] = hist.get(op, 0) + n
        total += n
    if total:
        print("\nExecuted warp instructions by opcode (first launch, share of "
              f"{total}):\n")
        top = sorted(hist.items(), key=lambda kv: -kv[1])[:16]
        print(", ".join(f"{op} {100.0 * n / total:.1f}%" for op, n in top))


if __name__ == "__main__":
    main()
