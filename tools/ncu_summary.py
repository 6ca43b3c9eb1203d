#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the few numbers the roofline discussion needs.
usage: ncu_summary.py report.ncu-rep [kernel-regex] > profiles/xxx.md   (runs `ncu -i` locally; no GPU needed)"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed.sum", "thread instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "ISSUE SLOTS busy % (1 inst/clk/SMSP peak)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (FP32 + IMAD) %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (int/logic/select) %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (MUFU, conversions) %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor pipe % (must be 0)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % (ncu 'compute SOL')"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier (warps/issue)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: fixed-latency wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (MUFU/smem)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    print("# ncu summary of `%s`\n" % rep.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on`; one column per profiled launch.\n")
    sel = [r for r in body if pat is None or pat.search(r[ik])]
    print("| metric | unit | " + " | ".join(r[ik][:48] for r in sel) + " |")
    print("|---|---|" + "---|" * len(sel))
    for name, label in METRICS:
        if name in hdr:
            i = hdr.index(name)
            print("| %s (`%s`) | %s | %s |" % (label, name, units[i], " | ".join(r[i] for r in sel)))


if __name__ == "__main__":
    main()
