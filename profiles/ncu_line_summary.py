#!/usr/bin/env python
"""Attribute the executed instructions of one kernel (from `ncu --page source --csv`) to source lines and to the
inlined functions they come from, by joining the SASS rows with `nvdisasm -gi` line info of the same cubin.

usage:
  cuobjdump -xelf all cedec-2024-rt_b200/libcedecrt.so          # -> *.cubin
  nvdisasm -gi -c kernels_fast.sm_100a.cubin > fast.sass
  ncu -i X.ncu-rep --page source --csv --kernel-id ::regex:NAME:1 > k.csv
  profiles/ncu_line_summary.py k.csv fast.sass MANGLED_SUBSTRING [top_n]

Rows are matched by position (the n-th SASS instruction of the function), after checking that the opcode agrees.
Output: share of warp instructions, lanes per instruction and stall samples per (file:line) and per function
(file + enclosing routine guessed from the outermost 'inlined at' frame).
"""
import csv
import re
import sys
from collections import Counter, defaultdict


def parse_sass(path, key):
    """[(opcode, [frames...])] for the function whose .text label contains `key`; frames: innermost first"""
    out, active, frames = [], False, []
    pend = []
    for line in open(path, errors="ignore"):
        if line.startswith(".text."):
            active = key in line
            continue
        if not active:
            continue
        s = line.strip()
        if s.startswith("//## File"):
            m = re.findall(r'File "([^"]+)", line (\d+)', s)
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', s)
            if m:
                pend.append([(m[0][0], int(m[0][1]))] + [(a, int(b)) for a, b in inl])
            continue
        m = re.match(r"/\*([0-9a-f]+)\*/\s+(.*?);", s)
        if m:
            if pend:
                # consecutive //## lines form the inline chain: innermost first
                chain = []
                for p in pend:
                    chain.extend(p)
                frames = chain
                pend = []
            ins = m.group(2).strip()
            toks = ins.split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            out.append((op.split(".")[0], frames))
    return out


def main():
    kcsv, sass, key = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    rows = list(csv.reader(open(kcsv)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    ia, ie, it, isamp = (h.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
    data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[ie].isdigit()]
    ins = parse_sass(sass, key)
    if len(ins) != len(data):
        print("warning: %d SASS rows in the report vs %d in the disassembly (different build?)" % (len(data), len(ins)))
    n = min(len(ins), len(data))
    by_line, by_file = defaultdict(lambda: [0, 0, 0]), defaultdict(lambda: [0, 0, 0])
    tot = tot_s = 0
    bad = 0
    for k in range(n):
        r = data[k]
        toks = r[ia].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        if op != ins[k][0]:
            bad += 1
        e, t, s = int(r[ie]), int(r[it]), int(r[isamp])
        tot += e
        tot_s += s
        frames = ins[k][1] or [("?", 0)]
        inner = frames[0]
        key_line = "%s:%d" % (inner[0].split("/")[-1], inner[1])
        for d, kk in ((by_line, key_line),):
            d[kk][0] += e
            d[kk][1] += t
            d[kk][2] += s
        # "own code" frame: the innermost frame that lies in this repo's csrc
        own = next((f for f in frames if "/csrc/" in f[0]), inner)
        kf = "%s:%d" % (own[0].split("/")[-1], own[1])
        by_file[kf][0] += e
        by_file[kf][1] += t
        by_file[kf][2] += s
    if bad:
        print("warning: %d opcode mismatches while aligning" % bad)
    print("warp instructions: %d   samples: %d" % (tot, tot_s))
    print("--- by own source line (innermost frame inside csrc/)")
    print("%-28s %7s %9s %9s" % ("line", "inst %", "thr/inst", "samples %"))
    for kf, (e, t, s) in sorted(by_file.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %6.2f%% %9.1f %8.2f%%" % (kf, 100.0 * e / tot, t / max(e, 1), 100.0 * s / max(tot_s, 1)))
    # per file totals
    files = Counter()
    for kf, (e, t, s) in by_file.items():
        files[kf.split(":")[0]] += e
    print("--- by file:", {k: round(100.0 * v / tot, 1) for k, v in files.most_common()})


if __name__ == "__main__":
    main()
