"""1D partition + halo exchange: host-side plan logic on CPU (thread ranks and a real world_size-2 gloo
group), and on the GPU the partitioned trainer against the single-GPU Model (thread ranks sharing cuda:0)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from graphaibench_b200 import dist as gdist  # noqa: E402
from graphaibench_b200 import ops  # noqa: E402


def _vp(a):
    assert a.flags.c_contiguous
    return a.reshape(-1)  # a view: the oracle's ctypes signatures take flat ndarrays


def _plans(world, nv, rp64, ci, selfloop=False, device="cpu"):
    def fn(comm):
        rp, cols = gdist.rows_of_rank(rp64, ci, world, comm.rank)
        rp, cols = torch.from_numpy(rp), torch.from_numpy(cols.astype(np.int64))
        if selfloop:
            rp, cols = gdist.add_selfloop_rows(rp, cols, gdist.owner_range(nv, world, comm.rank)[1])
        return gdist.HaloPlan(comm, nv, rp, cols, device=device)
    return gdist.ThreadComm.run(world, fn)


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_plan_matches_reference_partition(small_graph, world):
    """masters ∪ halo of every rank == idx_map of the reference-exact partitioner (gai_partition1d_h, itself pinned to the
    reference's edgecut_induced_partition1D); local CSR maps back to the global rows edge for edge."""
    rp64, ci, n = small_graph["rowptr64"], small_graph["colidx"], small_graph["n"]
    plans = _plans(world, n, rp64, ci)
    seen = np.zeros(n, np.int64)
    for r, p in enumerate(plans):
        ref = ops.partition1d(rp64, ci, world, r)
        ids = np.sort(np.concatenate([p.master_gids.numpy(), p.halo_gids.numpy()]))
        assert np.array_equal(ids, ref["idx_map"].astype(np.int64))
        S, first, last = gdist.owner_range(n, world, r)
        assert np.array_equal(np.sort(p.master_gids.numpy()), np.arange(first, last))
        seen[p.master_gids.numpy()] += 1
        # halo blocks are grouped by owner in rank order and sorted inside
        h = p.halo_gids.numpy()
        assert np.all(np.diff(h) > 0)
        assert np.array_equal(np.bincount(h // S, minlength=world), np.array(p.recv_counts))
        # local CSR -> global
        gid = np.concatenate([p.master_gids.numpy(), h])
        lrp, lci = p.rowptr.numpy().astype(np.int64), p.colidx.numpy().astype(np.int64)
        assert len(lrp) == p.m + 1 and np.all(lrp[p.n_loc:] == lrp[p.n_loc])
        for lr in list(range(min(p.n_loc, 50))) + list(range(max(p.n_loc - 50, 0), p.n_loc)):
            g = gid[lr]
            assert np.array_equal(gid[lci[lrp[lr]:lrp[lr + 1]]], ci[rp64[g]:rp64[g + 1]].astype(np.int64))
        # interior rows reference masters only
        if p.n_int:
            assert lci[: lrp[p.n_int]].max(initial=-1) < p.n_loc
        # every boundary row has a halo neighbour
        for lr in range(p.n_int, min(p.n_int + 50, p.n_loc)):
            assert lci[lrp[lr]:lrp[lr + 1]].max() >= p.n_loc
        # global degrees (masters and halo)
        assert np.array_equal(p.degree.numpy(), np.diff(rp64)[gid])
    assert np.all(seen == 1)
    # what rank q sends to rank r is exactly r's halo block of q
    for r, p in enumerate(plans):
        off = 0
        for q, pq in enumerate(plans):
            so = sum(pq.send_counts[:r])
            sent = pq.master_gids.numpy()[pq.send_ids.numpy()[so:so + pq.send_counts[r]]]
            assert np.array_equal(sent, p.halo_gids.numpy()[off:off + p.recv_counts[q]])
            off += p.recv_counts[q]


def test_selfloop_rows_match_oracle(small_graph, liborc):
    rp64, ci, n = small_graph["rowptr64"], small_graph["colidx"], small_graph["n"]
    rpo, cio = np.zeros(n + 1, np.uint32), np.zeros(len(ci) + n, np.uint32)
    liborc.orc_add_selfloop(n, _vp(small_graph["rowptr"]), _vp(ci), _vp(rpo), _vp(cio))
    for world in (1, 3):
        for r in range(world):
            _, first, last = gdist.owner_range(n, world, r)
            rp, cols = gdist.rows_of_rank(rp64, ci, world, r)
            rp2, c2 = gdist.add_selfloop_rows(torch.from_numpy(rp), torch.from_numpy(cols.astype(np.int64)), first)
            assert np.array_equal(rp2.numpy() + int(rpo[first]), rpo[first:last + 1].astype(np.int64))
            assert np.array_equal(c2.numpy(), cio[rpo[first]:rpo[last]].astype(np.int64))


def _aggregate_partitioned(comm, nv, rp64, ci, x, liborc, selfloop):
    """Halo exchange of x followed by the oracle's SpMM on the rank-local graph; returns (master gids, aggregated rows)."""
    rp, cols = gdist.rows_of_rank(rp64, ci, comm.world, comm.rank)
    rp, cols = torch.from_numpy(rp), torch.from_numpy(cols.astype(np.int64))
    if selfloop:
        rp, cols = gdist.add_selfloop_rows(rp, cols, gdist.owner_range(nv, comm.world, comm.rank)[1])
    p = gdist.HaloPlan(comm, nv, rp, cols)
    F = x.shape[1]
    B = torch.zeros(p.m, F)
    B[: p.n_loc] = torch.from_numpy(x)[p.master_gids]
    p.exchange_setup(B)
    lrp = p.rowptr.numpy().astype(np.uint32)
    lci = p.colidx.numpy().astype(np.uint32)
    Bn = np.ascontiguousarray(B.numpy())
    out = np.zeros((p.m, F), np.float32)
    if selfloop:
        ngcn, _ = p.norms()
        liborc.orc_spmm_gcn(p.m, _vp(lrp), _vp(lci), _vp(np.ascontiguousarray(ngcn.numpy())), F, _vp(Bn), _vp(out))
    else:
        liborc.orc_spmm_mean(p.m, _vp(lrp), _vp(lci), F, _vp(Bn), _vp(out), 0)
    return p.master_gids.numpy(), out[: p.n_loc]


def _aggregate_full(n, rp32, ci, x, liborc, selfloop):
    F = x.shape[1]
    out = np.zeros((n, F), np.float32)
    if selfloop:
        rpo, cio = np.zeros(n + 1, np.uint32), np.zeros(len(ci) + n, np.uint32)
        liborc.orc_add_selfloop(n, _vp(rp32), _vp(ci), _vp(rpo), _vp(cio))
        vd = np.zeros(n, np.float32)
        liborc.orc_vertex_norm(n, _vp(rpo), _vp(vd))
        liborc.orc_spmm_gcn(n, _vp(rpo), _vp(cio), _vp(vd), F, _vp(x), _vp(out))
    else:
        liborc.orc_spmm_mean(n, _vp(rp32), _vp(ci), F, _vp(x), _vp(out), 0)
    return out


@pytest.mark.parametrize("selfloop", [False, True])
@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_aggregation_bit_exact_cpu(small_graph, liborc, world, selfloop):
    """exchange + local aggregation == full-graph aggregation, bit for bit (edge order and global-degree norms survive the
    partition). GCN (self-loops, a_i*a_j) and SAGE mean."""
    rp64, ci, n = small_graph["rowptr64"], small_graph["colidx"], small_graph["n"]
    x = small_graph["x"][47]
    full = _aggregate_full(n, small_graph["rowptr"], ci, x, liborc, selfloop)
    res = gdist.ThreadComm.run(world, lambda comm: _aggregate_partitioned(comm, n, rp64, ci, x, liborc, selfloop))
    for gids, rows in res:
        assert np.array_equal(rows, full[gids])


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        z = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
        rp64, ci = z["sg_rowptr64"], z["sg_colidx"]
        n = len(rp64) - 1
        x = np.random.default_rng(11).standard_normal((n, 20), dtype=np.float32)
        lib = oracle.liborc()
        gids, rows = _aggregate_partitioned(gdist.TorchComm(), n, rp64, ci, x, lib, True)
        full = _aggregate_full(n, rp64.astype(np.uint32), ci, x, lib, True)
        ok = bool(np.array_equal(rows, full[gids]))
        # dW-style all-reduce through the same communicator
        t = torch.full((4,), float(rank + 1))
        gdist.TorchComm().all_reduce_sum(t)
        ok = ok and bool((t == sum(range(1, world + 1))).all())
        q.put((rank, ok, len(gids)))
    finally:
        dist.destroy_process_group()


def test_partitioned_aggregation_gloo_world2():
    """The same check through a real torch.distributed group (gloo, world_size 2, two processes)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r for r, _, _ in res) == [0, 1]
    assert all(ok for _, ok, _ in res)


# ---- GPU: partitioned trainer vs the single-GPU Model ---------------------------------------------------------------

def _train_partitioned(world, arch, g, feats, labels, split, dims, epochs, lr, overlap=True):
    rp64, ci, n = g["rowptr64"], g["colidx"], g["n"]

    def fn(comm):
        _, first, _ = gdist.owner_range(n, world, comm.rank)
        rp, cols = gdist.rows_of_rank(rp64, ci, world, comm.rank)
        rp, cols = torch.from_numpy(rp).cuda(), torch.from_numpy(cols.astype(np.int64)).cuda()
        if arch == "gcn":
            rp, cols = gdist.add_selfloop_rows(rp, cols, first)
        plan = gdist.HaloPlan(comm, n, rp, cols)
        gids = plan.master_gids
        mask = ((gids >= int(split[0])) & (gids < int(split[1]))).to(torch.uint8)
        m = gdist.DistGnn(arch, plan, torch.from_numpy(feats).cuda()[gids], torch.from_numpy(labels).cuda()[gids], mask,
                          int(split[1] - split[0]), dims, lr=lr, overlap=overlap)
        hist = [m.train_epoch() for _ in range(epochs)]
        torch.cuda.synchronize()
        first_aggr = None
        if dims[0] <= dims[1]:
            first_aggr = (gids.cpu().numpy(), m.A[0][: plan.n_loc, : dims[0]].cpu().numpy())
        return hist, [w.cpu().numpy() for w in m.W], [w.cpu().numpy() for w in m.Ws], first_aggr, (plan.n_int, plan.n_loc, plan.n_halo)

    if world == 1:
        return [fn(gdist.SelfComm())]
    return gdist.ThreadComm.run(world, fn)


@pytest.mark.gpu
@pytest.mark.parametrize("arch,world,hid", [("sage", 1, 128), ("sage", 2, 128), ("sage", 3, 64), ("gcn", 1, 128), ("gcn", 2, 64), ("gcn", 4, 128)])
def test_partitioned_training_matches_single_gpu(small_graph, arch, world, hid):
    """hid 128: layer 0 aggregates first (static input halo), layer 1 transforms first; hid 64: both layers transform first."""
    from conftest import require_cuda
    require_cuda()
    from graphaibench_b200 import datagen, model as gmodel
    g = small_graph
    n = g["n"]
    F, ncls, epochs, lr = 100, 47, 3, 0.01
    feats = g["x"][F]
    labels = datagen.labels(n, ncls, seed=3)
    split = datagen.split_ranges(n)
    ref = gmodel.GnnModel(arch, g["rowptr"], g["colidx"], feats, labels, split, hid, ncls, num_layers=2, lr=lr)
    ref_hist = []
    ref_aggr = None
    for e in range(epochs):
        ref_hist.append(ref.train_epoch())
        if e == 0 and arch == "sage" and hid >= F:
            ref_aggr = ref.get("in_temp1", 0).reshape(n, F)
    ref_W = [ref.get("W", l) for l in range(2)]
    ref_Ws = [ref.get("W_self", l) for l in range(2)] if arch == "sage" else []
    res = _train_partitioned(world, arch, g, feats, labels, split, [F, hid, ncls], epochs, lr)
    for hist, W, Ws, first_aggr, sizes in res:
        for (l0, a0), (l1, a1) in zip(ref_hist, hist):
            assert abs(l0 - l1) <= 1e-5 * max(1.0, abs(l0)), (ref_hist, hist)   # fp32 relative tolerance of north_star
            assert abs(a0 - a1) <= 1.5 / max(1, int(split[2]))                 # at most one near-tie argmax flips
        for l in range(2):
            assert np.abs(W[l].ravel() - ref_W[l]).max() <= 1e-5 * max(1.0, np.abs(ref_W[l]).max()) + 2e-5
            if arch == "sage":
                assert np.abs(Ws[l].ravel() - ref_Ws[l]).max() <= 1e-5 * max(1.0, np.abs(ref_Ws[l]).max()) + 2e-5
        if ref_aggr is not None:
            gids, rows = first_aggr
            assert np.array_equal(rows, ref_aggr[gids]), "partitioned aggregation must be bit-identical to the single-GPU one"
    if world > 1:
        assert sum(s[2] for _, _, _, _, s in res) > 0  # the graph really has cut edges


@pytest.mark.gpu
def test_partitioned_overlap_equals_serial(small_graph):
    """interior/boundary split with the side stream gives the same bits as exchange-then-aggregate."""
    from conftest import require_cuda
    require_cuda()
    from graphaibench_b200 import datagen
    g = small_graph
    n = g["n"]
    feats, labels, split = g["x"][100], datagen.labels(n, 47, seed=3), datagen.split_ranges(n)
    a = _train_partitioned(2, "sage", g, feats, labels, split, [100, 64, 47], 2, 0.01, overlap=True)
    b = _train_partitioned(2, "sage", g, feats, labels, split, [100, 64, 47], 2, 0.01, overlap=False)
    for (ha, Wa, Wsa, _, _), (hb, Wb, Wsb, _, _) in zip(a, b):
        assert ha == hb
        for x, y in zip(Wa + Wsa, Wb + Wsb):
            assert np.array_equal(x, y)


@pytest.mark.gpu
def test_row_segments_bit_exact(small_graph, liborc):
    """*_rows over registered segments (degree-ordered work lists, hub rows included) == the full-graph call."""
    from conftest import require_cuda
    require_cuda()
    g = small_graph
    n = g["n"]
    dg = ops.DeviceGraph(g["rowptr"], g["colidx"])
    cut = n // 3
    dg.set_row_segments([(0, cut), (cut, n)])
    for F in (47, 100, 256):
        x = torch.from_numpy(g["x"][F]).cuda()
        full = ops.spmm_mean(dg, x)
        out = torch.full_like(full, float("nan"))
        ops.spmm_mean(dg, x, out=out, rows=(0, cut))
        ops.spmm_mean(dg, x, out=out, rows=(cut, n))
        assert torch.equal(out, full)
        # padded leading dimension, F % 4 != 0: gathered directly (no staging copy), same bits
        ld = (F + 3) // 4 * 4 + 4
        xp = torch.randn(n, ld, device="cuda")
        xp[:, :F] = x
        out2 = torch.empty_like(full)
        ops.spmm_mean(dg, xp[:, :F], out=out2)
        assert torch.equal(out2, full)
