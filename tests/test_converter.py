"""CSR construction. The reference's text converter (src/converters/converter.cc) produced tests/golden/converter.json
for the .mtx inputs under tests/golden/mtx (two of them byte-identical copies of the files the reference ships); the numpy restatement
(oracle/convert.py) is pinned to those on CPU, the device construction (csrc/convert.cu through host/gai_converter.cpp and the gpu_converter
binary) on the GPU — bit-exact files — and against the restatement on large random COO inputs; add_selfloop on the device against the
host routine that is itself pinned to the reference (tests/test_abi.py)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT, require_cuda

GOLD = json.load(open(os.path.join(GOLDEN_DIR, "converter.json")))
MTX = os.path.join(GOLDEN_DIR, "mtx")


def sha_bytes(b):
    return hashlib.sha256(b).hexdigest()


def test_fixture_inputs_intact_and_reference_files_identical():
    for name, g in GOLD.items():
        data = open(os.path.join(MTX, name), "rb").read()
        assert sha_bytes(data) == g["input"], name
        ref = os.path.join("/root/reference/inputs", name)
        if os.path.exists(ref):
            assert open(ref, "rb").read() == data


@pytest.mark.parametrize("name", sorted(GOLD))
def test_restatement_matches_reference_converter(name):
    from oracle import convert
    nv, s, d, sym = convert.read_mtx_pairs(os.path.join(MTX, name))
    rp, ci = convert.coo_to_csr(nv, s, d, sym)
    assert (nv, len(ci)) == (GOLD[name]["nv"], GOLD[name]["ne"])
    assert sha_bytes(rp.tobytes()) == GOLD[name]["vertex_bin"] and sha_bytes(ci.tobytes()) == GOLD[name]["edge_bin"]


def test_live_reference_converter_matches_goldens(tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_convert")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/ref_convert not built (build container only)")
    for name, g in GOLD.items():
        subprocess.run([ref, "mtx", os.path.join(MTX, name), str(tmp_path / "g")], check=True, stdout=subprocess.DEVNULL)
        assert sha_bytes(open(tmp_path / "g.vertex.bin", "rb").read()) == g["vertex_bin"]
        assert sha_bytes(open(tmp_path / "g.edge.bin", "rb").read()) == g["edge_bin"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_converter_binary_is_bit_exact(name, tmp_path):
    require_cuda()
    from graphaibench_b200 import build
    build.build_all()
    out = subprocess.run([os.path.join(ROOT, "graphaibench_b200", "gpu_converter"), "mtx", os.path.join(MTX, name), str(tmp_path / "g")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert sha_bytes(open(tmp_path / "g.vertex.bin", "rb").read()) == GOLD[name]["vertex_bin"]
    assert sha_bytes(open(tmp_path / "g.edge.bin", "rb").read()) == GOLD[name]["edge_bin"]
    meta = open(tmp_path / "g.meta.txt").read().split()
    assert (int(meta[0]), int(meta[1])) == (GOLD[name]["nv"], GOLD[name]["ne"])


@pytest.mark.gpu
@pytest.mark.parametrize("nv,n,sym", [(1, 0, 1), (5, 3, 1), (70000, 3_000_000, 1), (70000, 3_000_000, 0), (1 << 20, 20_000_000, 1)])
def test_device_coo_to_csr_against_restatement(nv, n, sym):
    """Large random COO inputs (duplicates, self-loops, ids beyond nv): device keys / radix sort / unique / offsets against numpy, bit for bit."""
    require_cuda()
    import torch
    from graphaibench_b200 import ops
    from oracle import convert
    rng = np.random.default_rng(nv + n)
    s = rng.integers(0, nv + (2 if n else 0), n).astype(np.uint32)   # a few ids are out of range on purpose
    d = (s + rng.geometric(0.01, n)).astype(np.uint32) % np.uint32(nv + (2 if n else 0)) if n else s
    if n:
        s[: n // 50] = d[: n // 50]      # self-loops
        s[n // 2: n // 2 + n // 20] = s[: n // 20]; d[n // 2: n // 2 + n // 20] = d[: n // 20]   # duplicates
    rp_ref, ci_ref = convert.coo_to_csr(nv, s, d, bool(sym))
    rp, ci = ops.coo_to_csr(nv, torch.from_numpy(s.astype(np.int32)).cuda(), torch.from_numpy(d.astype(np.int32)).cuda(), symmetrize=bool(sym))
    assert np.array_equal(rp.cpu().numpy(), rp_ref)
    assert np.array_equal(ci.cpu().numpy().view(np.uint32), ci_ref)


@pytest.mark.gpu
def test_device_add_selfloop_matches_host(small_graph, cora):
    require_cuda()
    import torch
    from graphaibench_b200 import ops
    for rp, ci in ((small_graph["rowptr"], small_graph["colidx"]), (cora["rowptr"], cora["colidx"])):
        want_rp, want_ci = ops.add_selfloop(rp, ci)   # bit-exact vs the reference (tests/test_abi.py)
        got_rp, got_ci = ops.add_selfloop_device(torch.from_numpy(rp.astype(np.int32)).cuda(), torch.from_numpy(ci.astype(np.int32)).cuda())
        assert np.array_equal(got_rp.cpu().numpy().view(np.uint32), want_rp) and np.array_equal(got_ci.cpu().numpy().view(np.uint32), want_ci)
    # a block of rows with a global id offset (one rank of a 1D partition): row r gains the id first + r
    rp, ci = small_graph["rowptr"], small_graph["colidx"]
    full_rp, full_ci = ops.add_selfloop(rp, ci)
    first, last = 400, 900
    rows_rp = (rp[first:last + 1] - rp[first]).astype(np.int32)
    rows_ci = ci[rp[first]:rp[last]].astype(np.int32)
    got_rp, got_ci = ops.add_selfloop_device(torch.from_numpy(rows_rp).cuda(), torch.from_numpy(rows_ci).cuda(), first_id=first)
    assert np.array_equal(got_ci.cpu().numpy().view(np.uint32), full_ci[full_rp[first]:full_rp[last]])


@pytest.mark.gpu
def test_benchmark_graphs_are_built_by_the_device_converter(monkeypatch):
    """datagen.rmat_csr_torch builds its CSR with gai_coo_to_csr; the torch construction it replaced (sort + bincount + cumsum) must give
    the same arrays from the same pairs, whole graph and one rank's rows alike."""
    require_cuda()
    import torch
    from graphaibench_b200 import datagen
    for kw in (dict(), dict(permute=True, rows=(30000, 70000))):
        a = datagen.rmat_csr_torch(100000, 2_000_000, seed=3, device="cuda", **kw)
        monkeypatch.setenv("GAI_DATAGEN_TORCH", "1")
        b = datagen.rmat_csr_torch(100000, 2_000_000, seed=3, device="cuda", **kw)
        monkeypatch.delenv("GAI_DATAGEN_TORCH")
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
