#!/usr/bin/env python
"""One rank's aggregation of a partitioned run, on ONE GPU: builds rank `--rank` of the `--world` x configs[1] graph exactly as
bench.py --gpus N does (same R-MAT stream, same relabelling, same [masters | halo] numbering), fills masters and halo rows with random
features and times the mean aggregation of the master rows (no exchange, no peers). Prints the row statistics that decide the hub path
(longest rows, hub count) next to the times, so that the hub-row pipeline can be tuned without an N-GPU box."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--feat", default="100,47")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from graphaibench_b200 import ops
    sh = bench.make_shard(1, a.world, a.rank, "cuda")
    first, last = sh["first"], sh["last"]
    rp, ci = sh["rowptr"].to(torch.int64), sh["colidx"].to(torch.int64)
    n = last - first
    remote = (ci < first) | (ci >= last)
    halo = torch.unique(ci[remote])                      # ascending global ids
    loc = torch.where(remote, n + torch.searchsorted(halo, ci), ci - first).to(torch.int32)
    m = n + halo.numel()
    rp_all = torch.cat([rp, rp[-1].repeat(halo.numel())]).to(torch.int32)
    deg = (rp[1:] - rp[:-1])
    top = torch.topk(deg, 5).values.tolist()
    g = ops.DeviceGraph(rp_all.contiguous(), loc.contiguous(), device_arrays=True)
    g.set_row_segments([(0, n)])
    out = {"world": a.world, "rank": a.rank, "masters": n, "halo": int(halo.numel()), "edges": int(ci.numel()), "longest_rows": top, "n_hub": g.n_hub}
    for F in [int(x) for x in a.feat.split(",")]:
        pitch = (F + 3) // 4 * 4
        x = torch.randn(m, pitch, device="cuda")[:, :F]
        o = torch.empty(m, pitch, device="cuda")[:, :F]
        for _ in range(2):
            ops.spmm_mean(g, x, out=o, rows=(0, n))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            ops.spmm_mean(g, x, out=o, rows=(0, n))
        e1.record()
        torch.cuda.synchronize()
        out[f"mean_F{F}_ms"] = round(e0.elapsed_time(e1) / a.reps, 4)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
