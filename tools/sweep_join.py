#!/usr/bin/env python
"""Joins `tools/spmm_sweep.py --ncu` points with the per-launch counters of the ncu run that wrapped it:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmm_rows --csv \
        --log-file gpurun_out/sweep_ncu.csv python tools/spmm_sweep.py --ncu ... > gpurun_out/sweep_points.jsonl
    python tools/sweep_join.py gpurun_out/sweep_points.jsonl gpurun_out/sweep_ncu.csv [timed.jsonl] > profiles/r2_spmm_sweep.jsonl

Each point launched the aggregation kernel twice (warm-up, measured); the second launch's DRAM bytes become the point's `dram_bytes`.
If a third file is given (the same sweep run WITHOUT ncu, CUDA-event times), its `ms` replaces the under-ncu time and the three
bandwidths of SURVEY.md 8d are recomputed from it: gather_GBps, dram_GBps, compulsory_GBps."""
import csv, json, sys


def main():
    pts = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
    rows = [r for r in csv.reader(l for l in open(sys.argv[2]) if l.startswith('"'))]
    hdr = rows[0]
    iname, imetric, ival, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        launches.setdefault(int(r[iid]), {})[r[imetric]] = float(r[ival].replace(",", ""))
    ids = sorted(launches)
    assert len(ids) == 2 * len(pts), (len(ids), len(pts))
    timed = {}
    if len(sys.argv) > 3:
        for l in open(sys.argv[3]):
            if l.startswith("{"):
                d = json.loads(l)
                timed[(d["nv"], d["avg_deg"], d["F"], d["pitch"], d["mode"])] = d
    for k, p in enumerate(pts):
        m = launches[ids[2 * k + 1]]
        dram = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        t = timed.get((p["nv"], p["avg_deg"], p["F"], p["pitch"], p["mode"]))
        ms = t["ms"] if t else p["ms"]
        nv, nnz, F = p["nv"], p["nnz"], p["F"]
        b_gather = 4.0 * (nnz * F + nv * F + nnz + (nv + 1) + nv)
        b_comp = 4.0 * (2 * nv * F + nnz + 2 * nv + 1)
        out = dict(p, ms=ms, timed_without_profiler=bool(t), dram_bytes=dram, gather_GBps=round(b_gather / ms / 1e6, 1), dram_GBps=round(dram / ms / 1e6, 1),
                   compulsory_GBps=round(b_comp / ms / 1e6, 1), dram_frac_of_hbm_peak=round(dram / ms / 1e6 / p["hbm_peak_GBps"], 3),
                   dram_over_compulsory=round(dram / b_comp, 2))
        out.pop("under_ncu", None)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
