#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full --import-source on` capture of the search kernel) into the
short text file kept under profiles/: key raw metrics + the source lines with the most stall samples.

usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [> profiles/X.txt]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "smsp__sass_inst_executed_op_global_st.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__inst_executed_op_global_atom.sum", "smsp__inst_executed_op_shared_atom.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of {rep}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"\n## kernel: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:90s} {r[i]:>18s} {units[i]}")
    for view, title in (("cuda", "CUDA source lines"), ("sass", "SASS instructions")):
        src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", view))))
        hi = next((i for i, r in enumerate(src) if r and r[0] in ("Address", "#", "Line")), None)
        if hi is None:
            continue
        h = src[hi]
        if "Source" not in h or "# Samples" not in h:
            continue
        i_src, i_samp = h.index("Source"), h.index("# Samples")
        i_exec = h.index("Instructions Executed") if "Instructions Executed" in h else None
        i_thr = h.index("Avg. Threads Executed") if "Avg. Threads Executed" in h else None
        tot, tot_exec, per = 0, 0, []
        for r in src[hi + 1:]:
            try:
                s = int(float(r[i_samp]))
            except (ValueError, IndexError):
                continue
            ex = r[i_exec] if i_exec is not None else ""
            try:
                tot_exec += int(float(ex))
            except ValueError:
                pass
            tot += s
            per.append((s, r[i_src].strip(), ex, r[i_thr] if i_thr is not None else "", r[0]))
        per.sort(reverse=True)
        print(f"\n## top {title} by stall samples (total samples {tot}, warp instructions executed {tot_exec})")
        for s, text, ex, thr, loc in per[:45]:
            print(f"{s:9d} {100.0 * s / max(tot, 1):5.1f}%  exec={ex:>11s} thr={thr:>5s}  {loc[-6:]:>6s} {text[:120]}")


if __name__ == "__main__":
    main()
