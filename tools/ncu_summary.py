#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) for one kernel: headline metrics, stall mix, opcode mix,
hot SASS regions.  usage: ncu_summary.py report.ncu-rep [region_size]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 150
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_uniform.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print("%-82s %-10s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); ts = max(1, sum(int(r[isamp]) for r in data))
print("total warp instr", tot, "sass", len(data))


def opname(r):
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]


ops, samp = collections.Counter(), collections.Counter()
for r in data:
    ops[opname(r)] += int(r[ia]); samp[opname(r)] += int(r[isamp])
for op, c in ops.most_common(22):
    print("%-10s %10d %5.1f%%   samples %5.1f%%" % (op, c, 100 * c / tot, 100 * samp[op] / ts))
print("--- regions of %d sass" % B)
for i in range(0, len(data), B):
    blk = data[i:i + B]; c = sum(int(r[ia]) for r in blk); s = sum(int(r[isamp]) for r in blk)
    if c > tot * 0.015 or s > ts * 0.03:
        top = collections.Counter()
        for r in blk:
            top[opname(r)] += int(r[ia])
        print("sass %5d-%5d  instr %5.1f%% samples %5.1f%%  %s" % (i, i + B, 100 * c / tot, 100 * s / ts, top.most_common(6)))
