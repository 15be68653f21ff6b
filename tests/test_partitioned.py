"""1D-partitioned training behind the C++ API (host/gai_dist.h, gai_graph.cpp partition_rows, gai_model.cpp init_partitioned; csrc/peers.cu).

CPU: the integer part of the partition against the reference partitioner (PartitionedGraph::edgecut_induced_partition1D,
src/partitioner/graph_partition.cc:128-178, through the bit-exact gai_partition1d_h / the reference build), a world-size-2 gloo run of the
bootstrap all-gather + partition consistency, and the oracle's aggregation over partitioned rows (bit-exact with the unpartitioned one).
GPU: `world` ranks as host threads of one process (the whole peer-memory path on however many devices are visible) against the
single-GPU Model<L>."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, require_cuda


@pytest.fixture(scope="module")
def gm():
    from graphaibench_b200 import build
    build.build_all()
    from graphaibench_b200 import model
    return model


def rows_of(rp64, ci, first, last):
    rp = np.asarray(rp64[first:last + 1], np.int64)
    return rp - rp[0], np.asarray(ci[rp[0]:rp[-1]])


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("graph", ["cora", "small"])
def test_partition_rows_against_reference_partitioner(gm, cora, small_graph, world, graph):
    """The reference numbers masters and halo together in ascending global id (idx_map, masters at [local_begin, local_end)); this
    repository keeps the same two id lists with the masters first. Mapping one numbering onto the other must reproduce the reference's
    induced CSR on the master rows, edge for edge."""
    from graphaibench_b200 import ops
    rp64, ci = (cora["rowptr64"], cora["colidx"]) if graph == "cora" else (small_graph["rowptr64"], small_graph["colidx"])
    nv = len(rp64) - 1
    for rank in range(world):
        S, first, last = gm.owner_range(nv, world, rank)
        rrp, rci = rows_of(rp64, ci, first, last)
        rp, lci, halo = gm.partition_rows(world, rank, nv, rrp, rci)
        ref = ops.partition1d(rp64, ci, world, rank)
        lb, le, idx = ref["local_begin"], ref["local_end"], ref["idx_map"]
        n_loc = last - first
        assert le - lb == n_loc and np.array_equal(idx[lb:le], np.arange(first, last, dtype=np.uint32))
        assert np.array_equal(halo, np.concatenate([idx[:lb], idx[le:]]))          # same halo set, same (ascending) order
        assert np.array_equal(rp, rrp.astype(np.uint32))                            # master rows are complete: lengths untouched
        # reference local id j -> ours: masters j - lb; lower halo n_loc + j; upper halo n_loc + (j - n_loc)
        remap = np.where((np.arange(len(idx)) >= lb) & (np.arange(len(idx)) < le), np.arange(len(idx)) - lb,
                         np.where(np.arange(len(idx)) < lb, n_loc + np.arange(len(idx)), np.arange(len(idx)))).astype(np.uint32)
        sub_rp, sub_ci = ref["rowptr"], ref["colidx"]
        want = remap[sub_ci[sub_rp[lb]:sub_rp[le]]] if n_loc else np.zeros(0, np.uint32)
        assert np.array_equal(lci, want)


def test_selfloop_rows_then_partition_equals_partition_of_selflooped_graph(gm, small_graph):
    """GCN: add_selfloop on a rank's rows (global id first + r enters row r at its sorted place) commutes with slicing."""
    from graphaibench_b200 import ops
    rp64, ci = small_graph["rowptr64"], small_graph["colidx"]
    nv = len(rp64) - 1
    frp, fci = ops.add_selfloop(small_graph["rowptr"], ci)   # bit-exact vs the reference (tests/test_abi.py)
    for world in (2, 3):
        for rank in range(world):
            S, first, last = gm.owner_range(nv, world, rank)
            a = gm.partition_rows(world, rank, nv, *rows_of(rp64, ci, first, last), selfloops=True)
            b = gm.partition_rows(world, rank, nv, *rows_of(frp.astype(np.int64), fci, first, last))
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


def test_partitioned_aggregation_is_bit_exact_on_the_oracle(gm, small_graph, liborc):
    """Each rank aggregates its master rows over [masters | halo] with the local CSR: the rows it produces must be the bits of the
    unpartitioned aggregation (same edge order inside a row, global degrees)."""
    rp64, ci, n = small_graph["rowptr64"], small_graph["colidx"], small_graph["n"]
    x = small_graph["x"][16]
    full = np.zeros_like(x)
    liborc.orc_spmm_mean(n, small_graph["rowptr"], ci, 16, x.reshape(-1), full.reshape(-1), 0)
    for world in (2, 3):
        for rank in range(world):
            S, first, last = gm.owner_range(n, world, rank)
            rp, lci, halo = gm.partition_rows(world, rank, n, *rows_of(rp64, ci, first, last))
            m = (last - first) + len(halo)
            xl = np.concatenate([x[first:last], x[halo.astype(np.int64)]])          # what the halo pull delivers
            rp_ext = np.concatenate([rp, np.full(len(halo), rp[-1], np.uint32)])
            out = np.zeros((m, 16), np.float32)
            liborc.orc_spmm_mean(m, rp_ext, lci, 16, xl.reshape(-1), out.reshape(-1), 0)   # forward mean: 1/deg_i of the (complete) master row
            assert np.array_equal(out[: last - first], full[first:last])


_GLOO_WORKER = r"""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
from graphaibench_b200 import model as gm, datagen
nv = 3001
rp64, ci = datagen.rmat_csr(nv, 40000, seed=5)
S, first, last = gm.owner_range(nv, world, rank)
rp = np.asarray(rp64[first:last + 1], np.int64); rows = (rp - rp[0], ci[rp[0]:rp[-1]])
lrp, lci, halo = gm.partition_rows(world, rank, nv, *rows)
# the bootstrap all-gather the peer group uses (gai_allgather_fn), driven through its ctypes trampoline exactly as libgai_b200 calls it
cb = gm.torch_allgather_callback()
mine = np.zeros(4, np.uint64); mine[:] = (rank, last - first, len(halo), int(halo.astype(np.uint64).sum()))
allv = np.zeros(4 * world, np.uint64)
cb(None, mine.ctypes.data, mine.nbytes, allv.ctypes.data)
allv = allv.reshape(world, 4)
assert [int(r[0]) for r in allv] == list(range(world))
assert int(allv[:, 1].sum()) == nv
# every halo id is a master of the rank the ownership rule names, and never of this rank
owners = halo // S
assert (owners != rank).all() and (owners < world).all()
# halo lists travel too: rank q's needs, seen from here, are all inside [first, last) when it names this rank as owner
counts = np.zeros(world, np.int64); counts[:] = 0
need = torch.zeros(world, dtype=torch.int64); 
for q in range(world): need[q] = int((owners == q).sum())
got = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
dist.all_gather(got, need)
deg_local = np.diff(lrp.astype(np.int64))
tot = torch.tensor([int(deg_local.sum())]); dist.all_reduce(tot)
assert int(tot) == len(ci)
print("ok", rank, int(allv[rank, 2]))
"""


def test_bootstrap_and_partition_world2_gloo(tmp_path):
    """world_size 2 over gloo on CPU: the bootstrap all-gather callback and the per-rank partition agree across processes."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", str(port)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-3000:]
        assert out.strip().startswith("ok")


@pytest.mark.gpu
@pytest.mark.parametrize("arch,world,dims,layers", [("sage", 2, (100, 128, 47), 2), ("sage", 3, (36, 64, 7), 3), ("gcn", 2, (70, 32, 5), 3),
                                                    ("gcn", 4, (128, 256, 172), 3)])
def test_partitioned_training_matches_single_gpu(gm, arch, world, dims, layers):
    """`world` ranks (host threads; peer registration, halo pulls, weight-gradient and statistics combination in csrc/peers.cu) against the
    single-GPU Model<L> on the same graph: aggregated rows are bit-identical, so losses differ only through the order in which the ranks'
    partial weight gradients are added (fp32, 1e-5)."""
    require_cuda()
    from graphaibench_b200 import datagen
    nv, F, hid, ncls = 9000, dims[0], dims[1], dims[2]
    rp64, ci = datagen.rmat_csr(nv, 140000, seed=51)
    feats = datagen.features(nv, F, seed=52)
    labels = np.random.default_rng(53).integers(0, ncls, nv).astype(np.uint8)
    split = datagen.split_ranges(nv)
    epochs = 4
    single = gm.GnnModel(arch, rp64.astype(np.uint32), ci, feats, labels, split, hid, ncls, num_layers=layers, lr=0.01)
    ref_losses, ref_accs = zip(*[single.train_epoch() for _ in range(epochs)])
    ref_test = single.evaluate("test")
    got = gm.train_partitioned_inprocess(arch, world, rp64, ci, feats, labels, split, hid, ncls, num_layers=layers, lr=0.01, epochs=epochs)
    np.testing.assert_allclose(got["losses"], np.array(ref_losses, np.float32), rtol=2e-5)
    np.testing.assert_allclose(got["accs"], np.array(ref_accs, np.float32), atol=5e-4)
    assert abs(got["test_acc"] - ref_test) <= 2e-3
    w_single = np.concatenate([np.concatenate([single.get("W", l)] + ([single.get("W_self", l)] if arch == "sage" else [])) for l in range(layers)])
    # Adam divides by sqrt(v): a weight whose gradient is ~0 can move by a fraction of lr per step on a last-bit difference of that
    # gradient. The bulk of the weights must agree tightly; the few amplified ones stay within a few steps' worth of lr.
    d = np.abs(got["weights"] - w_single) / np.abs(w_single).max()
    assert np.quantile(d, 0.9) <= 5e-4 and np.quantile(d, 0.99) <= 2e-3, (float(np.quantile(d, 0.9)), float(np.quantile(d, 0.99)))
    assert d.max() <= 4 * epochs * 0.01 / np.abs(w_single).max(), float(d.max())
    assert int(got["halo"][:, 0].sum()) == nv and (got["halo"][:, 2] > 0).all() == (world > 1)


@pytest.mark.gpu
def test_pipelined_halo_exchange_is_bit_identical_to_one_piece(gm, monkeypatch):
    """Column-block pipelining of the halo exchange (Graph::halo_exchange_begin: blocks cross NVLink on a pull stream while the blocks
    already here are aggregated) changes the schedule, not one bit of the result: aggregation is independent per feature column."""
    require_cuda()
    from graphaibench_b200 import datagen
    nv, F, hid, ncls = 7000, 128, 256, 172
    rp64, ci = datagen.rmat_csr(nv, 110000, seed=61)
    feats = datagen.features(nv, F, seed=62)
    labels = np.random.default_rng(63).integers(0, ncls, nv).astype(np.uint8)
    split = datagen.split_ranges(nv)
    runs = {}
    for blocks in ("1", "2", "4"):
        monkeypatch.setenv("GAI_HALO_BLOCKS", blocks)
        runs[blocks] = gm.train_partitioned_inprocess("gcn", 3, rp64, ci, feats, labels, split, hid, ncls, num_layers=3, lr=0.01, epochs=3)
    for blocks in ("2", "4"):
        assert np.array_equal(runs[blocks]["losses"], runs["1"]["losses"])
        assert np.array_equal(runs[blocks]["weights"], runs["1"]["weights"])
