#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share of device time."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui, gi, bi = (hdr.index(c) for c in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = r[ki].split("(")[0][-70:]
    a = agg.setdefault((name, r[gi], r[bi]), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(data)} launches, {tot/1e3:.2f} ms total device time (serialised, cold cache)")
print(f"{'total ms':>10} {'calls':>6} {'avg us':>9} {'share':>6}  kernel  [grid x block]")
for (k, g, b), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]/1e3:10.3f} {a[0]:6d} {a[1]/a[0]:9.1f} {a[1]/tot*100:5.1f}%  {k}  [{g} x {b}]")
