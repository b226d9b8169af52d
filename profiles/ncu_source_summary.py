#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output for one kernel: opcode mix weighted by executed warp
instructions, SIMT efficiency per opcode, sample share and the warp-stall breakdown.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K | profiles/ncu_source_summary.py"""
import csv
import sys
from collections import Counter

rows = [r for r in csv.reader(sys.stdin)]
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
h = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[h.index("Instructions Executed")].isdigit()]
ia, ie, it, isamp = (h.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
tot = sum(int(r[ie]) for r in data)
print("kernel:", rows[0][1][:100] if len(rows[0]) > 1 else "?")
print("warp instructions executed: %d   thread instructions: %d   avg active threads: %.2f   SASS lines: %d" % (
    tot, sum(int(r[it]) for r in data), sum(int(r[it]) for r in data) / max(tot, 1), len(data)))
c, ct, cs = Counter(), Counter(), Counter()
for r in data:
    toks = r[ia].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    c[op] += int(r[ie])
    ct[op] += int(r[it])
    cs[op] += int(r[isamp])
ns = sum(cs.values())
print("%-10s %8s %10s %9s" % ("opcode", "inst %", "thr/inst", "samples %"))
for op, n in c.most_common(24):
    print("%-10s %7.1f%% %10.1f %8.1f%%" % (op, 100.0 * n / tot, ct[op] / max(n, 1), 100.0 * cs[op] / max(ns, 1)))
st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tots = {h[i]: sum(int(r[i]) for r in data) for i in st}
s = sum(tots.values())
print("stalls (all samples):", {k: round(100.0 * v / s, 1) for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]})
