#!/usr/bin/env python
"""DRAM traffic per unit of work from an `ncu --set full` capture that covers EVERY launch of a kernel in a short run.

  ncu_traffic.py TAG  ->  reads gpurun_out/TAG_{emb,search}_raw.csv (ncu --page raw --csv) and gpurun_out/TAG_plain.log (the same
  command without ncu: its `stats {...}` line gives the evaluations / searches of exactly those launches), writes
  profiles/r02_traffic.json = {kernel: {dram_bytes_per_unit, dram_bytes, units, launches, inst_per_unit, lts_bytes_per_unit, source}}
which bench.py multiplies by its own units per launch for `roofline.traffic`."""
import ast
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = {}
stats = None
for line in open(os.path.join(ROOT, "gpurun_out", tag + "_plain.log")):
    if line.startswith("stats "):
        stats = ast.literal_eval(line[6:].strip())
for name, kernel, unit_key in [("emb", "k_emb3", "emb_requests"), ("search", "k_search", "search_crimes")]:
    path = os.path.join(ROOT, "gpurun_out", "%s_%s_raw.csv" % (tag, name))
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units_row = rows[0], rows[1]

    def col(metric):
        if metric not in hdr:
            return float("nan")
        i = hdr.index(metric)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units_row[i], 1.0)
        return sum(float(r[i].replace(",", "")) * scale for r in rows[2:] if kernel in r[hdr.index("Kernel Name")])
    n = sum(1 for r in rows[2:] if kernel in r[hdr.index("Kernel Name")])
    dram = col("dram__bytes_read.sum") + col("dram__bytes_write.sum")
    units = stats[unit_key]
    out[kernel] = {"dram_bytes_per_unit": dram / units, "dram_bytes": dram, "dram_bytes_read": col("dram__bytes_read.sum"),
                   "dram_bytes_write": col("dram__bytes_write.sum"), "units": units, "launches": n,
                   "inst_per_unit": col("smsp__inst_executed.sum") / units, "lts_bytes_per_unit": 32.0 * col("lts__t_sectors.sum") / units,
                   "gpu_time_s": col("gpu__time_duration.sum"),
                   "source": "profiles/%s_%s_raw.csv: ncu --set full --clock-control none over all %d launches of %s in "
                             "`tools/profile_run.py` (see profiles/README.md), %d units" % (tag, name, n, kernel, units)}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
