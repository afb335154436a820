#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics of the
captured kernel plus the hottest source lines.  Usage:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import collections
import csv
import io
import linecache
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for k, row in enumerate(rows[2:]):
        name = row[hdr.index("Kernel Name")]
        print(f"== launch {k}: {name}")
        for h, u, v in zip(hdr, units, row):
            if h in KEYS:
                print(f"  {h:82s} {v} {u}")
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur, hd = None, None
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1]
            continue
        if "Instructions Executed" in r:
            hd = r
            li, ii, si, ti = (r.index("Line No"), r.index("Instructions Executed"), r.index("# Samples"),
                              r.index("Thread Instructions Executed"))
            continue
        if hd and len(r) == len(hd) and r[li].isdigit():
            try:
                a = agg[(cur, int(r[li]))]
                a[0] += int(r[ii]); a[1] += int(r[si]); a[2] += int(r[ti])
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print(f"== hottest source lines (share of warp instructions executed; {tot} total, {tots} stall samples)")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        path = f if os.path.exists(f) else os.path.join(ROOT, "robigo-luculenta_b200", "csrc", os.path.basename(f))
        text = linecache.getline(path, l).strip()[:80]
        print(f"  {100 * v[0] / tot:6.2f}% inst {100 * v[1] / tots:6.2f}% samples  lanes {v[2] / max(v[0], 1):5.1f}  "
              f"{os.path.basename(f)}:{l}  {text}")


if __name__ == "__main__":
    main()
