#!/usr/bin/env python
"""Epoch time + per-op breakdown of the other single-GPU configurations of BASELINE.json (bench.py measures configs[1]):

    c3   GAT 2-layer, hidden 256 (one head of width 256 = reference semantics, gat_layer.cpp), + l2norm + dense tail, on a synthetic
         Reddit-shaped R-MAT graph (232 965 vertices, ~114.6 M CSR edges, 602 features, 41 classes)
    c3h8 the same with 8 attention heads of 32 columns (GAI_GAT_HEADS=8: configs[2] as named; multi-head attention is an extension the
         reference does not have, restated in oracle/gnn_oracle.c orc_gat_*_heads)
    c4s  GCN 3-layer hidden 256 on ONE GPU's share of the papers100M shape at P = 8 (13.9 M vertices, ~202 M CSR edges, 128 features,
         172 classes) without the halo exchange: the per-GPU compute of configs[3]
    c2g  GCN 2-layer hidden 256 on the configs[1] graph (the GCN line of the headline metric)

One JSON line per configuration on stdout (same per-op fields as bench.py).  python tools/bench_configs.py c3 c4s [--scale S]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

CONFIGS = {
    "c3": dict(arch="gat", nv=232_965, nnz=114_600_000, feat=602, hid=256, ncls=41, layers=2, lr=0.01,
               name="GAT 2-layer hidden 256 (1 head) + l2norm + dense, Reddit-shaped R-MAT (BASELINE.json configs[2])"),
    "c3h8": dict(arch="gat", nv=232_965, nnz=114_600_000, feat=602, hid=256, ncls=41, layers=2, lr=0.01, heads=8,
                 name="GAT 2-layer hidden 256 = 8 heads x 32 + l2norm + dense, Reddit-shaped R-MAT (BASELINE.json configs[2] as named)"),
    "c4s": dict(arch="gcn", nv=13_882_495, nnz=202_000_000, feat=128, hid=256, ncls=172, layers=3, lr=0.01,
                name="GCN 3-layer hidden 256, one GPU's 1/8 share of the papers100M shape, no halo (BASELINE.json configs[3], per-GPU compute)"),
    "c2g": dict(arch="gcn", nv=2_449_029, nnz=62_000_000, feat=100, hid=256, ncls=47, layers=2, lr=0.01,
                name="GCN 2-layer hidden 256, products-shaped R-MAT (configs[1] graph)"),
    "c2s3": dict(arch="sage", nv=2_449_029, nnz=62_000_000, feat=100, hid=256, ncls=47, layers=3, lr=0.01,
                 name="GraphSAGE 3-layer hidden 256, products-shaped R-MAT (configs[1] graph, one more hidden layer)"),
}


def run(key, scale, steps, warmup):
    import torch
    from graphaibench_b200 import _abi, datagen, model as gmodel
    c = CONFIGS[key]
    if c.get("heads", 1) > 1:
        os.environ["GAI_GAT_HEADS"] = str(c["heads"])   # read by GAT_Aggregator::init when the model is built
    nv, nnz = c["nv"] // scale, c["nnz"] // scale
    dev = "cuda"
    rp, ci = datagen.rmat_csr_torch(nv, nnz, seed=1, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(2)
    feats = torch.randn(nv, c["feat"], generator=g, device=dev, dtype=torch.float32).cpu().numpy()
    g.manual_seed(3)
    labels = torch.randint(0, c["ncls"], (nv,), generator=g, device=dev, dtype=torch.int64).to(torch.uint8).cpu().numpy()
    rowptr, colidx = rp.cpu().numpy().astype(np.uint32), ci.cpu().numpy().astype(np.uint32)
    real_nnz = int(colidx.size)
    del rp, ci
    torch.cuda.empty_cache()
    stream = torch.cuda.Stream()
    with bench.quiet_stdout():
        m = gmodel.GnnModel(c["arch"], rowptr, colidx, feats, labels, datagen.split_ranges(nv), c["hid"], c["ncls"], num_layers=c["layers"],
                            lr=c["lr"], stream=stream.cuda_stream)
    torch.cuda.synchronize()
    L = _abi.lib()
    for _ in range(warmup):
        loss, acc = m.train_epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.gai_launch_count()
    e0.record(stream)
    for _ in range(steps):
        loss, acc = m.train_epoch()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = L.gai_launch_count() - l0
    gmodel.profile_enable(True)
    n_prof = 2
    for _ in range(n_prof):
        m.train_epoch()
    prof = gmodel.profile_collect()
    gmodel.profile_enable(False)
    roof, breakdown, per_shape = bench.roofline_from_profile(prof, bench.load_peaks(), n_prof)
    edges = real_nnz + (nv if c["arch"] != "sage" else 0)  # GCN / GAT train on the self-looped graph (net.cpp:96)
    print(json.dumps({"config": key, "workload": c["name"], "arch": c["arch"], "vertices": nv, "csr_edges": real_nnz, "edges_trained": edges,
                      "features": c["feat"], "hidden": c["hid"], "heads": c.get("heads", 1), "classes": c["ncls"], "layers": c["layers"], "scale_div": scale, "steps": steps,
                      "warmup": warmup, "epoch_ms": ms, "Medges_per_s": edges / ms / 1e3, "gpu_launches": int(launches), "roofline": roof,
                      "breakdown_ms_per_step": breakdown, "ops": per_shape[:16], "final": {"train_loss": loss, "train_acc": acc}}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    if len(a.configs) > 1:  # one process per configuration: the host mirror never frees layer buffers (as the reference)
        import subprocess
        for k in a.configs:
            subprocess.run([sys.executable, os.path.abspath(__file__), k, "--scale", str(a.scale), "--steps", str(a.steps), "--warmup", str(a.warmup)])
    else:
        t0 = time.time()
        run(a.configs[0], a.scale, a.steps, a.warmup)
        print(f"[{a.configs[0]}] wall {time.time() - t0:.1f} s", file=sys.stderr)
