#!/usr/bin/env python
"""Reads an .ncu-rep (one kernel) and prints (1) headline metrics, (2) stall reasons, (3) a segment table:
runs of SASS instructions with the same execution count, with their share of executed instructions and the
source lines (from -lineinfo, via nvdisasm) that dominate each run.
Usage: python tools/ncu_segments.py REPORT.ncu-rep CUBIN_TU_NAME(e.g. dxb_k_bc7) UNITS_PER_LAUNCH"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, tu, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
d = dict(zip(raw[0], raw[2]))
for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
    print("%-70s %s" % (k, d.get(k, "?")))
print("warp instructions per unit: %.1f" % (float(d["smsp__inst_executed.sum"]) / units))
for k, v in sorted(d.items()):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(v) > 0.1:
                print("  stall %-40s %.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], float(v)))
        except ValueError:
            pass

sass = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
hdr = sass[1]
ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
prof = [(r[1].strip(), int(r[ie]), int(r[ns])) for r in sass[2:] if len(r) > ie]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "directxtex_b200", "_lib", "libdxtex_b200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith(tu + ".")][0]
dis = run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)]).split("\n")
funcs, ins, cur = [], None, None          # one instruction list per function of the cubin; the profiled one = the list of equal length
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        ins = []
        funcs.append(ins)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and ins is not None:
        ins.append(cur)
ins = ([f for f in funcs if len(f) == len(prof)] or [max(funcs, key=len) if funcs else []])[0]
if len(ins) != len(prof):
    print("warning: disassembly (%d) and profile (%d) lengths differ (library rebuilt since the capture?)" % (len(ins), len(prof)))
    ins = (ins + [None] * len(prof))[:len(prof)]
tot = sum(p[1] for p in prof)
ts = max(1, sum(p[2] for p in prof))
warps = max(p[1] for p in prof[:20])
print("static SASS instructions: %d;  executed per first-instruction execution: %.1f" % (len(prof), tot / warps))
i = 0
while i < len(prof):
    j = i
    while j + 1 < len(prof) and abs(prof[j + 1][1] - prof[i][1]) <= 0.02 * max(prof[i][1], 1):
        j += 1
    n = sum(p[1] for p in prof[i:j + 1])
    s = sum(p[2] for p in prof[i:j + 1])
    if n / tot > 0.006:
        c = collections.Counter(ins[k][1] if ins[k] else 0 for k in range(i, j + 1))
        print("idx %5d len %4d exec/warp %5.2f instr %5.1f%% samples %5.1f%%  lines %s" % (
            i, j - i + 1, prof[i][1] / warps, 100 * n / tot, 100 * s / ts, " ".join("%dx%d" % kv for kv in c.most_common(5))))
    i = j + 1
