#!/usr/bin/env python
"""Pretty-print a bench.py JSON line (file argument): epoch time, per-bucket breakdown, per-op table."""
import json, sys
for l in open(sys.argv[1]):
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    print(f"ms_per_step {d['ms_per_step']:.3f}  value {d['value']:.1f} {d['unit']}  e2e {d['e2e'].get('ms_per_step', 0):.2f} ms  launches {d.get('gpu_launches')}")
    print("breakdown", d.get("breakdown_ms_per_step"))
    print("roofline", d.get("roofline"))
    for o in d.get("ops", []):
        print("   ", o)
    if "cpu_baseline" in d:
        print("cpu", d["cpu_baseline"])
