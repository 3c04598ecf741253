#!/usr/bin/env python
"""Writes the text summary of one .ncu-rep that profiles/ keeps: the capture command, the headline metrics (raw page)
and the stall / segment table of tools/ncu_segments.py.
Usage: python tools/ncu_summary.py REPORT.ncu-rep TU_NAME UNITS "capture command / notes" > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

rep, tu, units, note = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
KEYS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread launch__shared_mem_per_block_dynamic
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active
sm__inst_executed_pipe_tensor.sum sm__throughput.avg.pct_of_peak_sustained_elapsed
dram__bytes_read.sum dram__bytes_write.sum dram__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
sm__cycles_elapsed.avg.per_second""".split()
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
names, unitsrow, vals = raw[0], raw[1], raw[2]
d = {n: (u, v) for n, u, v in zip(names, unitsrow, vals)}
print("# " + note)
print("# kernel: " + d.get("Kernel Name", ("", "?"))[1])
for k in KEYS:
    if k in d:
        print("%-72s %-14s %s" % (k, d[k][0], d[k][1]))
print()
print(subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_segments.py"), rep, tu, units], capture_output=True, text=True).stdout)
