#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` log: one line per profiled launch (test tooling)."""
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
i_id, i_k, i_m, i_v = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
out, kern = {}, {}
for r in rows[1:]:
    out.setdefault(int(r[i_id]), {})[r[i_m]] = float(r[i_v].replace(",", ""))
    kern[int(r[i_id])] = r[i_k][:60]
for k, v in sorted(out.items()):
    print(k, kern[k], " ".join(f"{m.split('.')[0].split('__')[-1]}={x:.4g}" for m, x in v.items()))
