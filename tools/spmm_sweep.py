#!/usr/bin/env python
"""Standalone SpMM aggregation sweep (BASELINE.json configs[4], SURVEY.md §8d "C5"): R-MAT power-law graphs, GCN-normalised and
mean aggregation at several feature widths, timed with CUDA events; prints one JSON line per point with the three
bandwidth figures of §8d (gather model B_spmm/t, compulsory bound/t) against the measured HBM peak.

  python tools/spmm_sweep.py [--nv 1000000,4000000] [--deg 16,64] [--feat 16,32,64,128,256,512] [--reps 5] [--modes gcn,mean]
                             [--pitch aligned|dense|both] [--ncu]

--pitch   row pitch of the gathered matrix: `pad4` = the layer classes' rows padded to 4 floats (host/gai_layers.h row_pitch: 47 -> 48),
          `aligned` = 128-byte-aligned rows (47 -> 64, 100 -> 128; measured and rejected), `dense` = pitch F (a caller's dense matrix).
--ncu     exactly one warm-up and one measured launch per point, in the order printed: run the tool under
          `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:spmm_` and join the per-launch DRAM
          bytes to the points with tools/sweep_join.py (the DRAM column of SURVEY.md 8d's triple; times under ncu are not bench values).
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", default="1000000")
    ap.add_argument("--deg", default="16,64")
    ap.add_argument("--feat", default="16,32,64,128,256,512")
    ap.add_argument("--modes", default="gcn,mean")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--max-gb", type=float, default=150.0)
    ap.add_argument("--max-edges", type=float, default=2.2e9, help="skip graphs beyond this many CSR edges (R-MAT generation memory)")
    ap.add_argument("--pitch", default="pad4", choices=["pad4", "aligned", "dense", "both"])
    ap.add_argument("--ncu", action="store_true")
    args = ap.parse_args()
    import torch
    from graphaibench_b200 import datagen, ops
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    for nv in [int(x) for x in args.nv.split(",")]:
        for deg in [int(x) for x in args.deg.split(",")]:
            if nv * deg >= (1 << 32) - nv or nv * deg > args.max_edges:   # 32-bit edge offsets (gai_csr_create); generator memory
                continue
            rp, ci = datagen.rmat_csr_torch(nv, nv * deg, seed=1, device="cuda")
            rp32 = rp.to(torch.int32); ci32 = ci.to(torch.int32)
            nnz = int(ci32.numel())
            g = ops.DeviceGraph(rp32, ci32, device_arrays=True)
            for F, pitch_kind in [(int(x), k) for x in args.feat.split(",") for k in (("aligned", "dense") if args.pitch == "both" else (args.pitch,))]:
                pitch = F if pitch_kind == "dense" else ((F + 3) // 4 * 4 if pitch_kind == "pad4" else
                                                         (next(p for p in (4, 8, 16, 32) if F <= p) if F <= 32 else (F + 31) // 32 * 32))
                if pitch_kind == "dense" and args.pitch == "both" and pitch == F and F % 32 == 0:
                    continue  # identical layouts
                if 4.0 * (2 * nv * pitch + nnz) / 1e9 > args.max_gb:
                    continue
                xb = torch.zeros(nv, pitch, device="cuda")
                xb[:, :F] = torch.randn(nv, F, device="cuda")
                x = xb[:, :F]
                out = torch.zeros(nv, pitch, device="cuda")[:, :F]
                for mode in args.modes.split(","):
                    fn = (lambda: ops.spmm_gcn(g, x, out=out)) if mode == "gcn" else (lambda: ops.spmm_mean(g, x, out=out, transposed=(mode == "meanT")))
                    reps = 1 if args.ncu else args.reps
                    for _ in range(1 if args.ncu else 3):
                        fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(reps):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / reps
                    b_gather = 4.0 * (nnz * F + nv * F + nnz + (nv + 1) + nv)
                    b_comp = 4.0 * (2 * nv * F + nnz + 2 * nv + 1)
                    print(json.dumps({"nv": nv, "nnz": nnz, "avg_deg": round(nnz / nv, 1), "F": F, "pitch": pitch, "mode": mode, "ms": round(ms, 4), "n_hub": g.n_hub,
                                      "under_ncu": bool(args.ncu),
                                      "gather_GBps": round(b_gather / ms / 1e6, 1), "compulsory_GBps": round(b_comp / ms / 1e6, 1),
                                      "edges_x_feats_per_s": round(nnz * F / ms * 1e3, 0), "frac_of_hbm_peak": round(b_gather / ms / 1e6 / peak, 3),
                                      "hbm_peak_GBps": peak}), flush=True)
                del x, out, xb
            del g, rp, ci, rp32, ci32
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
