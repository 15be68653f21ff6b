#!/usr/bin/env python
"""Partitioned full-graph training on N GPUs of one box, through the C++ partitioned Model (host/gai_model.cpp init_partitioned over
csrc/peers.cu), for the two BASELINE.json configurations that are run 1D-partitioned:

  --config c4   configs[3]: GCN 3-layer hidden 256 on an ogbn-papers100M-shaped R-MAT graph (111 M vertices, 1.6 B CSR edges, 128 features,
                172 classes), divided by --scale (the largest graph the run holds is named in the output)
  --config c2   configs[1]'s shape per GPU (what bench.py --gpus N runs), for a same-id-order N = 1 point

WEAK scaling: the graph has N x (the per-GPU share) vertices and edges; vertex ids are randomly relabelled at EVERY N, N = 1 included, so
that the points of a scaling curve differ in the number of ranks only (the reference's contiguous ownership rule,
graph_partition.cc:131-140, needs the relabelling to balance an R-MAT graph's edges). Launch under torchrun for N > 1:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_dist.py --config c4 --scale 2

Prints one JSON line on rank 0: epoch ms (CUDA events on the launching stream, max over ranks), edges/s, per-rank masters / halo / edges,
halo exchanges and bytes per epoch with the NVLink rate they ran at, and the per-op breakdown."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {
    "c4": dict(arch="gcn", nv=111_059_956, nnz=1_616_000_000, feat=128, hid=256, ncls=172, layers=3, lr=0.01,
               name="GCN 3-layer hidden 256, ogbn-papers100M-shaped R-MAT graph (BASELINE.json configs[3])"),
    "c2": dict(arch="sage", nv=2_449_029 * 8, nnz=62_000_000 * 8, feat=100, hid=256, ncls=47, layers=2, lr=0.01,
               name="GraphSAGE-mean 2-layer hidden 256, 8 x the ogbn-products shape (BASELINE.json configs[1] per GPU at N = 8)"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4", choices=sorted(SHAPES))
    ap.add_argument("--scale", type=float, default=1.0, help="divide the 8-GPU graph by this (per-GPU share = shape / 8 / scale)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from graphaibench_b200 import _abi, datagen, model as gmodel
    L = _abi.lib()
    S = SHAPES[args.config]
    per_gpu_nv = int(S["nv"] / 8 / args.scale)
    per_gpu_nnz = int(S["nnz"] / 8 / args.scale)
    nv, nnz = per_gpu_nv * world, per_gpu_nnz * world
    _, first, last = gmodel.owner_range(nv, world, rank)
    t0 = time.time()
    rp, ci = datagen.rmat_csr_torch(nv, nnz, seed=1, device="cuda", permute=True, rows=(first, last))
    rows_rp = rp.cpu().numpy().astype(np.int64)
    rows_ci = ci.to(torch.int32).cpu().numpy().view(np.uint32)
    del rp, ci
    g = torch.Generator(device="cuda"); g.manual_seed(2 + 1000 * rank)
    feats = torch.randn(last - first, S["feat"], generator=g, device="cuda", dtype=torch.float32).cpu().numpy()
    g.manual_seed(3 + 1000 * rank)
    labels = torch.randint(0, S["ncls"], (last - first,), generator=g, device="cuda", dtype=torch.int64).to(torch.uint8).cpu().numpy()
    split = datagen.split_ranges(nv)
    torch.cuda.empty_cache()
    t_gen = time.time() - t0
    stream = torch.cuda.Stream()
    cb = gmodel.torch_allgather_callback(device=torch.device("cuda", local)) if world > 1 else None
    t0 = time.time()
    m = gmodel.DistGnnModel(S["arch"], rank, world, cb, nv, rows_rp, rows_ci, feats, labels, split, S["hid"], S["ncls"], num_layers=S["layers"],
                            lr=S["lr"], stream=stream.cuda_stream)
    torch.cuda.synchronize()
    t_setup = time.time() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        m.train_epoch()
    barrier()
    h0 = m.halo_stats()
    l0 = L.gai_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        loss, acc = m.train_epoch()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    h1 = m.halo_stats()
    launches = L.gai_launch_count() - l0
    gmodel.profile_enable(True)
    for _ in range(2):
        m.train_epoch()
    prof = gmodel.profile_collect()
    gmodel.profile_enable(False)
    m.check()
    by = {}
    for r in prof:
        by[r["bucket"]] = by.get(r["bucket"], 0.0) + r["ms"] / 2
    sizes = torch.tensor([h1["masters"], h1["halo"], len(rows_ci), torch.cuda.max_memory_allocated() / 2**30], device="cuda", dtype=torch.float64)
    gathered = [torch.empty_like(sizes) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, sizes)
    else:
        gathered = [sizes]
    if rank != 0:
        return
    free, total = torch.cuda.mem_get_info()
    per_rank = [dict(masters=int(t[0]), halo=int(t[1]), edges=int(t[2])) for t in gathered]
    edges = sum(r["edges"] for r in per_rank)
    ex = (h1["exchanges"] - h0["exchanges"]) / args.steps
    exb = (h1["bytes"] - h0["bytes"]) / args.steps
    halo_ms = by.get("HALO", 0.0)
    print(json.dumps({
        "config": args.config, "workload": S["name"], "n_gpus": world, "scale_div": args.scale, "vertices": nv, "csr_edges": edges,
        "epoch_ms": ms, "Medges_per_s": edges / ms / 1e3, "steps": args.steps, "warmup": args.warmup, "gpu_launches": int(launches),
        "ids": "randomly relabelled at every N (N = 1 included)", "per_rank": per_rank,
        "halo": {"exchanges_per_epoch": ex, "recv_bytes_per_epoch_rank0": exb, "ms_per_epoch_rank0_serialised": halo_ms,
                 "GBps_rank0": exb / max(halo_ms, 1e-9) / 1e6, "nvlink_measured_peer_copy_GBps_per_dir": 770.0,
                 "note": "each exchange = flag barrier + pull + flag barrier; the wait for the slowest rank is inside this time"},
        "breakdown_ms_per_epoch": {k: round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])},
        "device_mem_used_GiB_rank0": round((total - free) / 2**30, 1), "setup_s": {"graph": round(t_gen, 1), "model": round(t_setup, 1)},
        "final": {"train_loss": loss, "train_acc": acc}}), flush=True)


if __name__ == "__main__":
    main()
