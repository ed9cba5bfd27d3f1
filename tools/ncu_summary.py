#!/usr/bin/env python
"""Summarises an .ncu-rep for profiles/: key raw metrics per captured launch, the executed-instruction mix and the
source lines with the most stall samples.  usage: tools/ncu_summary.py REPORT [--top N]   (prints markdown)"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    head, units, rows = raw[0], raw[1], raw[2:]
    col = {h: i for i, h in enumerate(head)}
    print(f"Kernel: `{rows[0][col['Kernel Name']]}`\n")
    print("| metric | unit | " + " | ".join(f"launch {k + 1}" for k in range(len(rows))) + " |")
    print("|---|---|" + "---|" * len(rows))
    for k in KEYS:
        if k in col:
            print(f"| `{k}` | {units[col[k]]} | " + " | ".join(r[col[k]] for r in rows) + " |")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    # find header of the first kernel's table
    hdr = None
    tables = []
    for r in src:
        if r and r[0] in ("Address", "#") or (r and "Source" in r and "Address" in r):
            hdr = r
            tables.append([])
        elif hdr and r and len(r) == len(hdr):
            tables[-1].append(r)
    if not tables:
        return
    t = tables[0]
    c = {h: i for i, h in enumerate(hdr)}
    inst_col = c.get("# Instructions Executed") or c.get("Instructions Executed")
    samp_col = c.get("# Samples") if "# Samples" in c else c.get("Warp Stall Sampling (All Samples)")
    ops = defaultdict(int)
    total = 0
    for r in t:
        try:
            n = int(r[inst_col])
        except Exception:
            continue
        op = r[c["Source"]].split()[0] if r[c["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[c["Source"]].split()[1]
        ops[op.split(".")[0]] += n
        total += n
    print(f"\nExecuted warp instructions by opcode (first launch, {total} in all): " +
          ", ".join(f"{o} {100.0 * n / total:.1f}%" for o, n in sorted(ops.items(), key=lambda kv: -kv[1])[:22]))
    if samp_col is not None:
        ranked = sorted(t, key=lambda r: -int(r[samp_col] or 0))[:top]
        allsamp = sum(int(r[samp_col] or 0) for r in t)
        print(f"\nSASS instructions with the most stall samples (of {allsamp}):\n")
        print("| samples | share | address | instruction |")
        print("|---|---|---|---|")
        for r in ranked:
            print(f"| {r[samp_col]} | {100.0 * int(r[samp_col] or 0) / max(allsamp, 1):.1f}% | {r[c['Address']]} | `{r[c['Source']][:90]}` |")


if __name__ == "__main__":
    main()
