#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by (kernel, source line):
warp-stall samples and executed instructions.  Usage: ncu_lines.py dump.csv [kernel-substring] [top-n]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = csv.reader(open(path))
kern, fpath, hdr = None, None, None
agg = defaultdict(lambda: [0, 0, ""])  # (kernel, file, line) -> samples, instr, text
launch = defaultdict(int)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        kern = r[1]; launch[kern] += 1; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or kern is None:
        continue
    if r[2] == "-" and r[0].isdigit():  # a source line row
        try:
            s = int(r[hdr.index("# Samples")]); n = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        k = (kern, fpath, int(r[0]))
        agg[k][0] += s; agg[k][1] += n; agg[k][2] = r[1]
kernels = sorted({k[0] for k in agg})
for kn in kernels:
    if want not in kn:
        continue
    items = [(v[0], v[1], k[1], k[2], v[2]) for k, v in agg.items() if k[0] == kn]
    tot_s = sum(i[0] for i in items) or 1
    tot_n = sum(i[1] for i in items) or 1
    print("== %s: %d samples, %d warp-instructions" % (kn, tot_s, tot_n))
    for s, n, f, ln, txt in sorted(items, reverse=True)[:topn]:
        print("  %5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * s / tot_s, 100.0 * n / tot_n, f, ln, txt.strip()[:110]))
