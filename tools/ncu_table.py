#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` output into a compact per-launch table (CSV on stdout): duration, DRAM bytes and
throughput, L2 hit rate, occupancy, issue utilisation, tensor-pipe utilisation, registers, top stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, data = rows[0], rows[2:]
def col(name):
    return hdr.index(name) if name in hdr else None
want = [("kernel", "Kernel Name"), ("grid", "Grid Size"), ("block", "Block Size"), ("ms", "gpu__time_duration.sum"),
        ("dram_read_GB", "dram__bytes_read.sum"), ("dram_write_GB", "dram__bytes_write.sum"),
        ("dram_pct_of_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
        ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"), ("regs", "launch__registers_per_thread"),
        ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"), ("inst_executed", "smsp__inst_executed.sum")]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warp") and "stalled" in h and h.endswith("ratio")]
w = csv.writer(sys.stdout)
w.writerow([n for n, _ in want] + ["dram_GBps", "top_stalls"])
for r in data:
    out = []
    vals = {}
    for n, c in want:
        i = col(c)
        v = r[i] if i is not None and i < len(r) else ""
        if n == "kernel":
            v = v.split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("gai::", "")
        vals[n] = v
        out.append(v)
    try:
        gbps = (float(vals["dram_read_GB"]) + float(vals["dram_write_GB"])) / float(vals["ms"]) * 1e3
    except ValueError:
        gbps = float("nan")
    st = sorted(((float(r[i].replace(",", "")), hdr[i].replace("smsp__average_warp_latency_issue_stalled_", "").replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", "")) for i in stall_cols if r[i] not in ("", "n/a")), reverse=True)[:3]
    w.writerow(out + [f"{gbps:.0f}", " ".join(f"{n}={v:.1f}" for v, n in st)])
