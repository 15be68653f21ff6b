#!/usr/bin/env python
"""Top stall hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:80], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections[:1]:
    hdr, data = sec["hdr"], sec["data"]
    iS, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[iS]) for r in data)
    print(sec["name"], "| total samples", tot, "| instrs", len(data))
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    for r in data:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    top = sorted(enumerate(data), key=lambda t: -int(t[1][iS]))[:ntop]
    for idx, r in sorted(top):
        st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
        print(idx, r[iS], r[iex], r[isrc].strip()[:64], st)
