#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + SASS opcode mix per particle-step.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [particle_steps_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 1024 * 4096
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = [
    "Kernel Name", "Block Size", "Grid Size",
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
print("== raw metrics ==")
for h, u, v in zip(hdr, units, vals):
    if h in keep or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print("%-90s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cat, wf, samp = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    n = float(r[ix["Instructions Executed"]] or 0)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    cat[op] += n
    wf[op] += float(r[ix["L1 Wavefronts Shared"]] or 0)
    samp[op] += float(r[ix["# Samples"]] or 0)
tot = sum(cat.values())
print("== SASS mix: thread-instructions per particle-step (total %.1f) ==" % (tot * 32 / steps))
for op, n in cat.most_common(28):
    print("%-10s %7.2f   smem wavefronts/pstep %.3f   stall samples %5.1f%%" % (op, n * 32 / steps, wf[op] / steps, 100 * samp[op] / max(1, sum(samp.values()))))
