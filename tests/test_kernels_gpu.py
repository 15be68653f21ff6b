"""GPU parity tests (run with -m gpu on the B200 box): every kernel behind the C ABI against the CPU oracle on the same
seeded inputs and against the golden vectors generated from the reference. Bit-exact for index work and for the sparse
fp32 aggregation (same operation order and roundings as the reference CPU path); stated tolerances elsewhere."""
import ctypes as C

import numpy as np
import pytest

from conftest import require_cuda, sha

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north-star fp32 tolerance, applied norm-wise: max|a-b| <= REL_TOL * max|ref|


def close(a, ref, tol=REL_TOL):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(a - ref).max() / scale
    assert err <= tol, f"norm-wise relative error {err:.3e} > {tol}"


@pytest.fixture(scope="module")
def T():
    require_cuda()
    import torch
    from graphaibench_b200 import build
    build.build_cuda()
    return torch


@pytest.fixture(scope="module")
def ops(T):
    from graphaibench_b200 import ops
    return ops


def dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a)).cuda()


def test_csr_norms_bit_exact(T, ops, golden, small_graph):
    rp, ci = ops.add_selfloop(small_graph["rowptr"], small_graph["colidx"])
    g = ops.DeviceGraph(rp, ci)
    assert g.n_hub >= 1  # the fixture has a hub row (> 1024 neighbours)
    assert np.array_equal(g.vertex_norm().cpu().numpy(), golden["sg_loop_vdata"])
    drp, dci = g.csr()
    assert np.array_equal(drp.cpu().numpy().view(np.uint32), rp) and np.array_equal(dci.cpu().numpy().view(np.uint32), ci)


@pytest.mark.parametrize("F", [7, 16, 47, 100, 256])
def test_spmm_bit_exact_vs_reference_goldens(T, ops, golden, small_graph, F):
    x = dev(T, small_graph["x"][F])
    rp, ci = ops.add_selfloop(small_graph["rowptr"], small_graph["colidx"])
    g_loop = ops.DeviceGraph(rp, ci)
    g_raw = ops.DeviceGraph(small_graph["rowptr"], small_graph["colidx"])
    assert sha(ops.spmm_gcn(g_loop, x).cpu().numpy()) == str(golden[f"sg_gcn_{F}_sha"])
    assert sha(ops.spmm_mean(g_raw, x).cpu().numpy()) == str(golden[f"sg_mean_{F}_sha"])
    assert sha(ops.spmm_mean(g_raw, x, transposed=True).cpu().numpy()) == str(golden[f"sg_meanT_{F}_sha"])


@pytest.mark.parametrize("F", [1, 2, 5, 12, 36, 64, 130, 172, 512, 600, 1100])
def test_spmm_bit_exact_vs_oracle_widths(T, ops, liborc, small_graph, F):
    """All vector widths / lane-group shapes / column-block loops, incl. the hub-row kernel."""
    from oracle import model as om
    n = small_graph["n"]
    rng = np.random.default_rng(F)
    x = rng.standard_normal((n, F), dtype=np.float32)
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.add_selfloop(); g.compute_vertex_data()
    ref = np.zeros((n, F), np.float32)
    liborc.orc_spmm_gcn(n, g.rowptr, g.colidx, g.vdata, F, x.reshape(-1), ref.reshape(-1))
    dg = ops.DeviceGraph(g.rowptr, g.colidx)
    out = ops.spmm_gcn(dg, dev(T, x)).cpu().numpy()
    assert np.array_equal(out, ref)
    vals = rng.standard_normal(g.ne, dtype=np.float32)
    liborc.orc_spmm_edge(n, g.rowptr, g.colidx, vals, F, x.reshape(-1), ref.reshape(-1))
    out = ops.spmm_edge(dg, dev(T, vals), dev(T, x)).cpu().numpy()
    assert np.array_equal(out, ref)


def test_spmm_epilogue_ld_and_row_ranges(T, ops, liborc, small_graph):
    from oracle import model as om
    n, F = small_graph["n"], 100
    x = small_graph["x"][F]
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.compute_vertex_data()
    ref = np.zeros((n, F), np.float32)
    liborc.orc_spmm_mean(n, g.rowptr, g.colidx, F, x.reshape(-1), ref.reshape(-1), 0)
    dg = ops.DeviceGraph(g.rowptr, g.colidx)
    add = np.random.default_rng(3).standard_normal((n, F), dtype=np.float32)
    # fused (+addend, ReLU) epilogue == unfused reference sequence
    out = ops.spmm_mean(dg, dev(T, x), flags=ops.EPI_ADD | ops.EPI_RELU, addend=dev(T, add)).cpu().numpy()
    assert np.array_equal(out, np.maximum(ref + add, 0))
    # leading dimensions: input is a column slice of a wider buffer, output lands inside a wider buffer
    wide_in = T.zeros(n, 2 * F + 4, device="cuda"); wide_in[:, 4:4 + F] = dev(T, x)
    wide_out = T.full((n, 3 * F), -7.0, device="cuda")
    ops.spmm_mean(dg, wide_in[:, 4:4 + F], out=wide_out[:, F:2 * F])
    assert np.array_equal(wide_out[:, F:2 * F].cpu().numpy(), ref)
    assert float(wide_out[:, :F].max()) == -7.0 and float(wide_out[:, 2 * F:].min()) == -7.0
    # row ranges (1D partition: interior / boundary split) reproduce the full result, hub row included
    out = T.full((n, F), 5.0, device="cuda")
    ops.spmm_mean(dg, dev(T, x), out=out, rows=(0, 700))
    assert float(out[700:].min()) == 5.0
    ops.spmm_mean(dg, dev(T, x), out=out, rows=(700, n))
    assert np.array_equal(out.cpu().numpy(), ref)


def test_spmm_empty_and_degenerate(T, ops):
    g = ops.DeviceGraph(np.zeros(5, np.uint32), np.zeros(0, np.uint32))  # 4 isolated vertices
    out = ops.spmm_mean(g, T.ones(4, 8, device="cuda"))
    assert float(out.abs().max()) == 0.0
    g0 = ops.DeviceGraph(np.zeros(1, np.uint32), np.zeros(0, np.uint32))  # empty graph
    assert ops.spmm_gcn(g0, T.ones(0, 8, device="cuda")).shape == (0, 8)


def test_transpose_perm_bit_exact(T, ops, small_graph):
    from oracle import model as om
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.add_selfloop()
    dg = ops.DeviceGraph(g.rowptr, g.colidx)
    assert np.array_equal(dg.transpose_perm().cpu().numpy().view(np.uint32), g.transpose_perm())
    # asymmetric pattern is rejected, not silently mis-transposed
    from graphaibench_b200 import GaiError
    bad = ops.DeviceGraph(np.array([0, 1, 1], np.uint32), np.array([1], np.uint32))
    with pytest.raises(GaiError):
        bad.transpose_perm()


@pytest.mark.parametrize("shape", [(300, 40, 70, 0, 0), (257, 7, 16, 0, 0), (1433, 16, 2708, 1, 0), (2708, 1433, 16, 0, 1),
                                   (129, 65, 33, 1, 1), (5000, 256, 100, 0, 0), (100, 256, 5000, 1, 0)])
@pytest.mark.parametrize("mode", [1, 0])
def test_matmul_vs_oracle(T, ops, liborc, shape, mode):
    from graphaibench_b200._abi import lib
    x, y, z, ta, tb = shape
    rng = np.random.default_rng(x + y)
    A = rng.standard_normal((z, x) if ta else (x, z), dtype=np.float32)
    B = rng.standard_normal((y, z) if tb else (z, y), dtype=np.float32)
    C0 = rng.standard_normal((x, y), dtype=np.float32)
    lib().gai_set_gemm_mode(mode)
    try:
        for accum in (0, 1):
            ref = C0.copy()
            liborc.orc_gemm(x, y, z, A.reshape(-1), B.reshape(-1), ref.reshape(-1), ta, tb, accum)
            out = dev(T, C0.copy())
            ops.matmul(dev(T, A), dev(T, B), out=out, transA=bool(ta), transB=bool(tb), accum=bool(accum))
            close(out.cpu().numpy(), ref)
        ref = np.zeros((x, y), np.float32)
        liborc.orc_gemm(x, y, z, A.reshape(-1), B.reshape(-1), ref.reshape(-1), ta, tb, 0)
        out = ops.matmul(dev(T, A), dev(T, B), transA=bool(ta), transB=bool(tb), flags=ops.EPI_RELU)
        close(out.cpu().numpy(), np.maximum(ref, 0))
    finally:
        lib().gai_set_gemm_mode(0)


def test_relu_drelu_bit_exact(T, ops, liborc):
    rng = np.random.default_rng(1)
    for n in (1, 3, 1024, 100003):
        x = rng.standard_normal(n, dtype=np.float32); gr = rng.standard_normal(n, dtype=np.float32)
        r1, r2 = np.zeros(n, np.float32), np.zeros(n, np.float32)
        liborc.orc_relu(n, x, r1); liborc.orc_d_relu(n, gr, x, r2)
        assert np.array_equal(ops.relu(dev(T, x)).cpu().numpy(), r1)
        assert np.array_equal(ops.d_relu(dev(T, gr), dev(T, x)).cpu().numpy(), r2)
        xd = dev(T, x)
        ops.relu(xd, out=xd)  # in place, as the layers call it
        assert np.array_equal(xd.cpu().numpy(), r1)


def test_softmax_ce_vs_reference_golden(T, ops, golden):
    logits, labs, masks = golden["loss_logits"], golden["loss_labels"], golden["loss_masks"]
    nv, ncls = logits.shape
    probs = T.zeros(nv, ncls, device="cuda"); losses = T.zeros(nv, device="cuda"); grad = T.zeros(nv, ncls, device="cuda")
    dl, dlab, dm = dev(T, logits), dev(T, labs), dev(T, masks)
    ops.softmax_ce_forward(dl, dlab, dm, 5, 40, probs, losses)
    ops.softmax_ce_backward(probs, dlab, dm, 5, 40, grad)
    stats = ops.masked_loss_accuracy(dl, dlab, dm, 5, 40, losses).cpu().numpy()
    close(probs.cpu().numpy(), golden["loss_probs"], 2e-6)
    close(grad.cpu().numpy(), golden["loss_grad"], 2e-6)
    assert abs(stats[0] - golden["loss_value"]) <= 1e-5 * abs(golden["loss_value"])
    assert stats[1] == golden["loss_acc"] and stats[2] == 35
    assert float(grad[:5].abs().max()) == 0.0 and float(grad[40:].abs().max()) == 0.0  # untouched outside the range


def test_adam_bit_exact_vs_reference_golden(T, ops, golden):
    W = dev(T, golden["adam_W"].copy()); m = T.zeros_like(W); v = T.zeros_like(W)
    b1t, b2t = np.float32(0.9), np.float32(0.999)
    for s in range(golden["adam_grads"].shape[0]):
        ops.adam_update(dev(T, golden["adam_grads"][s]), W, m, v, 0.02, float(b1t), float(b2t))
        b1t, b2t = np.float32(b1t * np.float32(0.9)), np.float32(b2t * np.float32(0.999))
    assert np.array_equal(W.cpu().numpy(), golden["adam_W_after"])


def test_l2norm_vs_oracle(T, ops, liborc):
    rng = np.random.default_rng(2)
    n, dim = 513, 37
    x = rng.standard_normal((n, dim), dtype=np.float32); x[3] = 0  # a zero row exercises the 1e-12 clamp
    gin = rng.standard_normal((n, dim), dtype=np.float32)
    r1, r2 = np.zeros((n, dim), np.float32), np.zeros((n, dim), np.float32)
    liborc.orc_l2norm(n, dim, x.reshape(-1), r1.reshape(-1)); liborc.orc_d_l2norm(n, dim, x.reshape(-1), gin.reshape(-1), r2.reshape(-1))
    close(ops.l2norm(dev(T, x)).cpu().numpy(), r1)
    close(ops.d_l2norm(dev(T, x), dev(T, gin)).cpu().numpy(), r2)


def test_gat_forward_backward_vs_reference_golden(T, ops, golden, small_graph):
    rp, ci = ops.add_selfloop(small_graph["rowptr"], small_graph["colidx"])
    g = ops.DeviceGraph(rp, ci)
    z, gin = dev(T, golden["gat_z"]), dev(T, golden["gat_gin"])
    out, temp, norm = ops.gat_forward(g, z, dev(T, golden["gat_al"]), dev(T, golden["gat_ar"]))
    close(norm.cpu().numpy(), golden["gat_norm_scores"])
    close(out.cpu().numpy(), golden["gat_out"])
    dz, dal, dar, _ = ops.gat_backward(g, z, gin, temp, norm)
    close(dz.cpu().numpy(), golden["gat_gout"])
    close(dal.cpu().numpy(), golden["gat_dal"], 2e-5)  # long fp32 reductions in a different (deterministic) order
    close(dar.cpu().numpy(), golden["gat_dar"], 2e-5)
    # dz may alias z (the reference overwrites out_temp with dZ)
    z2 = z.clone()
    ops.gat_backward(g, z2, gin, temp, norm, dz=z2)
    assert np.array_equal(z2.cpu().numpy(), dz.cpu().numpy())


@pytest.mark.parametrize("F", [64, 256, 512])
def test_gat_wide_vs_oracle(T, ops, liborc, small_graph, F):
    from oracle import model as om
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.add_selfloop()
    n = g.nv
    rng = np.random.default_rng(F)
    z = rng.standard_normal((n, F), dtype=np.float32) * 0.3; gin = rng.standard_normal((n, F), dtype=np.float32)
    al = rng.standard_normal(F, dtype=np.float32) * 0.2; ar = rng.standard_normal(F, dtype=np.float32) * 0.2
    ts, sc, ns, nsg = (np.zeros(g.ne, np.float32) for _ in range(4))
    out, gout = np.zeros((n, F), np.float32), np.zeros((n, F), np.float32)
    dal, dar = np.zeros(F, np.float32), np.zeros(F, np.float32)
    liborc.orc_gat_forward(n, g.rowptr, g.colidx, F, al, ar, 0.2, z.reshape(-1), ts, sc, ns, out.reshape(-1))
    liborc.orc_gat_backward(n, g.rowptr, g.colidx, F, 0.2, z.reshape(-1), gin.reshape(-1), ts, ns, sc, nsg, dal, dar, gout.reshape(-1), 1)
    dg = ops.DeviceGraph(g.rowptr, g.colidx)
    o, temp, norm = ops.gat_forward(dg, dev(T, z), dev(T, al), dev(T, ar))
    close(o.cpu().numpy(), out); close(norm.cpu().numpy(), ns)
    dz, d_al, d_ar, _ = ops.gat_backward(dg, dev(T, z), dev(T, gin), temp, norm)
    close(dz.cpu().numpy(), gout); close(d_al.cpu().numpy(), dal, 5e-5); close(d_ar.cpu().numpy(), dar, 5e-5)


@pytest.mark.parametrize("F,H", [(32, 2), (64, 8), (256, 8), (256, 2), (128, 32), (512, 4)])
def test_gat_multi_head_vs_oracle(T, ops, liborc, small_graph, F, H):
    """Multi-head attention (extension: el/er, scores, row softmax, SDDMM and both aggregations per head; edge-major [nnz x H] score
    arrays) against its oracle restatement on the power-law test graph (hub row and isolated vertices included), 1e-5 norm-wise."""
    from oracle import model as om
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.add_selfloop()
    n = g.nv
    rng = np.random.default_rng(F * 37 + H)
    z = rng.standard_normal((n, F), dtype=np.float32) * 0.3; gin = rng.standard_normal((n, F), dtype=np.float32)
    al = rng.standard_normal(F, dtype=np.float32) * 0.2; ar = rng.standard_normal(F, dtype=np.float32) * 0.2
    ts, sc, ns, nsg = (np.zeros(g.ne * H, np.float32) for _ in range(4))
    out, gout = np.zeros((n, F), np.float32), np.zeros((n, F), np.float32)
    dal, dar = np.zeros(F, np.float32), np.zeros(F, np.float32)
    liborc.orc_gat_forward_heads(n, g.rowptr, g.colidx, F, H, al, ar, 0.2, z.reshape(-1), ts, sc, ns, out.reshape(-1))
    liborc.orc_gat_backward_heads(n, g.rowptr, g.colidx, F, H, 0.2, z.reshape(-1), gin.reshape(-1), ts, ns, sc, nsg, dal, dar, gout.reshape(-1), 1)
    dg = ops.DeviceGraph(g.rowptr, g.colidx)
    o, temp, norm = ops.gat_forward_heads(dg, dev(T, z), H, dev(T, al), dev(T, ar))
    close(norm.cpu().numpy(), ns); close(temp.cpu().numpy(), ts); close(o.cpu().numpy(), out)
    dz, d_al, d_ar, ds = ops.gat_backward_heads(dg, dev(T, z), H, dev(T, gin), temp, norm)
    close(dz.cpu().numpy(), gout); close(d_al.cpu().numpy(), dal, 5e-5); close(d_ar.cpu().numpy(), dar, 5e-5)
    # heads == 1 through the same entry points is the single-head path
    o1, t1, n1 = ops.gat_forward_heads(dg, dev(T, z), 1, dev(T, al), dev(T, ar))
    o0, t0, n0 = ops.gat_forward(dg, dev(T, z), dev(T, al), dev(T, ar))
    assert np.array_equal(o1.cpu().numpy(), o0.cpu().numpy()) and np.array_equal(n1.cpu().numpy(), n0.cpu().numpy())


def test_gather_rows(T, ops):
    rng = np.random.default_rng(4)
    src = rng.standard_normal((1000, 100), dtype=np.float32)
    ids = rng.integers(0, 1000, 333).astype(np.int32)
    out = ops.gather_rows(dev(T, ids), dev(T, src)).cpu().numpy()
    assert np.array_equal(out, src[ids])


@pytest.mark.parametrize("shape", [(5000, 256, 100, 0), (8192, 47, 256, 0), (6000, 256, 47, 1), (4096, 16, 1433, 0), (20000, 100, 256, 1),
                                   (4500, 7, 16, 0), (33000, 172, 128, 0)])
@pytest.mark.parametrize("pair", ["0", "1"])
def test_matmul_tcgen05_3xtf32_vs_oracle(T, ops, liborc, shape, pair, monkeypatch):
    """tcgen05/TMEM path (mode 2 = forced, 3xTF32): fp32-level accuracy (norm-wise 1e-5) on aligned and unaligned widths,
    ragged row counts, transposed weights, accumulate and ReLU epilogues. Mode 3 (single TF32 pass) is ~1e-3 by design.
    pair = 1: the opt-in cta_group::2 kernel (two CTAs of a cluster on one M = 256 MMA; outputs wider than 128 columns only), including
    odd row-block counts whose last pair has one out-of-bounds half."""
    from graphaibench_b200._abi import lib
    monkeypatch.setenv("GAI_TC_PAIR", pair)
    x, y, z, tb = shape
    if pair == "1" and y <= 128:
        pytest.skip("the pair kernel takes outputs wider than 128 columns")
    rng = np.random.default_rng(x + 3 * y + z)
    A = rng.standard_normal((x, z), dtype=np.float32)
    B = rng.standard_normal((y, z) if tb else (z, y), dtype=np.float32)
    C0 = rng.standard_normal((x, y), dtype=np.float32)
    ref = np.zeros((x, y), np.float32)
    liborc.orc_gemm(x, y, z, A.reshape(-1), B.reshape(-1), ref.reshape(-1), 0, tb, 0)
    try:
        lib().gai_set_gemm_mode(2)
        out = ops.matmul(dev(T, A), dev(T, B), transB=bool(tb))
        close(out.cpu().numpy(), ref)
        out = ops.matmul(dev(T, A), dev(T, B), transB=bool(tb), flags=ops.EPI_RELU)
        close(out.cpu().numpy(), np.maximum(ref, 0))
        acc = dev(T, C0.copy())
        ops.matmul(dev(T, A), dev(T, B), out=acc, transB=bool(tb), accum=True)
        ref2 = C0.copy()
        liborc.orc_gemm(x, y, z, A.reshape(-1), B.reshape(-1), ref2.reshape(-1), 0, tb, 1)
        close(acc.cpu().numpy(), ref2)
        lib().gai_set_gemm_mode(3)
        out = ops.matmul(dev(T, A), dev(T, B), transB=bool(tb))
        close(out.cpu().numpy(), ref, 3e-3)
        err1 = np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err1 > 1e-6, "single-pass TF32 should be visibly less accurate than 3xTF32 (is the tensor-core path running?)"
    finally:
        lib().gai_set_gemm_mode(0)


@pytest.mark.parametrize("shape", [(100, 256, 50000), (256, 47, 30011), (47, 256, 9000), (256, 256, 20000), (16, 7, 4100), (128, 172, 33333),
                                   (200, 130, 12345)])
def test_wgrad_tcgen05_3xtf32_vs_oracle(T, ops, liborc, shape):
    """dW = X^T·G on the tcgen05 MN-major split-over-rows kernel (mode 2 forced): every output tile shape (1 and 2 M tiles,
    unaligned widths -> padded staging, ragged row counts), accumulate; against the oracle GEMM and an fp64 product."""
    from graphaibench_b200._abi import lib
    kx, my, n = shape
    rng = np.random.default_rng(kx + 3 * my + n)
    X = rng.standard_normal((n, kx), dtype=np.float32)
    G = rng.standard_normal((n, my), dtype=np.float32)
    C0 = rng.standard_normal((kx, my), dtype=np.float32)
    ref = np.zeros((kx, my), np.float32)
    liborc.orc_gemm(kx, my, n, X.reshape(-1), G.reshape(-1), ref.reshape(-1), 1, 0, 0)
    ref64 = X.astype(np.float64).T @ G.astype(np.float64)
    try:
        lib().gai_set_gemm_mode(2)
        out = ops.matmul(dev(T, X), dev(T, G), transA=True).cpu().numpy()
        close(out, ref)
        close(out, ref64, 5e-6)
        acc = dev(T, C0.copy())
        ops.matmul(dev(T, X), dev(T, G), out=acc, transA=True, accum=True)
        close(acc.cpu().numpy(), ref64 + C0, 5e-6)
        lib().gai_set_gemm_mode(3)
        out1 = ops.matmul(dev(T, X), dev(T, G), transA=True).cpu().numpy()
        close(out1, ref64, 3e-3)
        assert np.abs(out1 - ref64).max() / np.abs(ref64).max() > 1e-6, "single-pass TF32 should be visibly less accurate"
    finally:
        lib().gai_set_gemm_mode(0)


@pytest.mark.parametrize("shape", [(20000, 256, 100, 100, 0), (9000, 256, 47, 47, 1), (4100, 64, 33, 8, 0), (12345, 172, 128, 40, 1),
                                   (2000, 16, 30, 20, 0)])
@pytest.mark.parametrize("padded", [False, True])
def test_matmul_kcat_vs_oracle(T, ops, liborc, shape, padded):
    """C = A1·op(B1) + A2·op(B2) in one pass (SAGE neighbour + self transforms, sage_layer.cpp:20-23,44-52): tensor-core path
    for x >= 4096 (two TMA maps, one TMEM accumulator), SIMT composition below; ReLU and d_ReLU-mask epilogues; operands in
    16-byte-padded row layouts (ld = width rounded up to 4) and in the reference's dense layout."""
    x, y, z1, z2, tb = shape
    rng = np.random.default_rng(x + y + z1 + z2)
    pad = (lambda w: (w + 3) // 4 * 4) if padded else (lambda w: w)

    def mat(r, c):
        full = np.zeros((r, pad(c)), np.float32)
        full[:, :c] = rng.standard_normal((r, c), dtype=np.float32)
        return full
    A1, A2 = mat(x, z1), mat(x, z2)
    B1 = rng.standard_normal((y, z1) if tb else (z1, y), dtype=np.float32)
    B2 = rng.standard_normal((y, z2) if tb else (z2, y), dtype=np.float32)
    M = mat(x, y)
    ref = np.zeros((x, y), np.float32)
    a1c, a2c = np.ascontiguousarray(A1[:, :z1]), np.ascontiguousarray(A2[:, :z2])
    liborc.orc_gemm(x, y, z1, a1c.reshape(-1), B1.reshape(-1), ref.reshape(-1), 0, tb, 0)
    liborc.orc_gemm(x, y, z2, a2c.reshape(-1), B2.reshape(-1), ref.reshape(-1), 0, tb, 1)
    dA1, dA2, dM = dev(T, A1)[:, :z1], dev(T, A2)[:, :z2], dev(T, M)[:, :y]
    out = T.zeros(x, pad(y), device="cuda")[:, :y]
    ops.matmul_kcat(dA1, dev(T, B1), dA2, dev(T, B2), out=out, transB=bool(tb))
    close(out.cpu().numpy(), ref)
    ops.matmul_kcat(dA1, dev(T, B1), dA2, dev(T, B2), out=out, transB=bool(tb), flags=ops.EPI_RELU)
    close(out.cpu().numpy(), np.maximum(ref, 0))
    ops.matmul_kcat(dA1, dev(T, B1), dA2, dev(T, B2), out=out, transB=bool(tb), flags=ops.EPI_MASK, mask=dM)
    close(out.cpu().numpy(), np.where(M[:, :y] > 0, ref, 0))
    if padded:
        # GAI_EPI_PADDED: the caller owns rows padded to 4 floats; the padding columns may be overwritten, but only with zeros
        full = T.full((x, pad(y)), 7.0, device="cuda")
        ops.matmul_kcat(dA1, dev(T, B1), dA2, dev(T, B2), out=full[:, :y], transB=bool(tb), flags=ops.EPI_RELU | ops.EPI_PADDED)
        close(full[:, :y].cpu().numpy(), np.maximum(ref, 0))
        tail = full[:, y:]
        assert bool(((tail == 7.0) | (tail == 0.0)).all())
        ops.matmul_kcat(dA1, dev(T, B1), dA2, dev(T, B2), out=full[:, :y], transB=bool(tb), flags=ops.EPI_MASK | ops.EPI_PADDED, mask=dM)
        close(full[:, :y].cpu().numpy(), np.where(M[:, :y] > 0, ref, 0))
        # single-operand masked input gradient (GCN / GAT layers) and the accumulating form on the same shapes
        ops.matmul_mask(dA1, dev(T, B1), dM, out=full[:, :y], transB=bool(tb), flags=ops.EPI_PADDED)
        ref1 = np.zeros((x, y), np.float32)
        liborc.orc_gemm(x, y, z1, a1c.reshape(-1), B1.reshape(-1), ref1.reshape(-1), 0, tb, 0)
        close(full[:, :y].cpu().numpy(), np.where(M[:, :y] > 0, ref1, 0))
        ops.matmul(dA1, dev(T, B1), out=full[:, :y], transB=bool(tb), flags=ops.EPI_PADDED)
        ops.matmul(dA2, dev(T, B2), out=full[:, :y], transB=bool(tb), accum=True, flags=ops.EPI_RELU | ops.EPI_PADDED)
        close(full[:, :y].cpu().numpy(), np.maximum(ref, 0))
        # sign bits: the ReLU epilogue writes one bit per activation, the masked input gradient of the layer above reads them back
        nw = (y + 31) // 32
        bits = T.zeros(x, nw, dtype=T.int32, device="cuda")
        ops.matmul_kcat(dA1, dev(T, B1), dA2, dev(T, B2), out=full[:, :y], transB=bool(tb), flags=ops.EPI_RELU | ops.EPI_PADDED, relu_bits=bits)
        act = full[:, :y].cpu().numpy().copy()
        close(act, np.maximum(ref, 0))
        want = np.zeros((x, nw * 32), bool); want[:, :y] = act > 0
        got = np.unpackbits(bits.cpu().numpy().view(np.uint8), axis=1, bitorder="little").astype(bool)
        assert np.array_equal(got, want), "sign bits differ from (activation > 0)"
        ops.matmul_kcat(dA1, dev(T, B1), dA2, dev(T, B2), out=full[:, :y], transB=bool(tb), flags=ops.EPI_MASK | ops.EPI_BITMASK | ops.EPI_PADDED, mask=bits)
        close(full[:, :y].cpu().numpy(), np.where(act > 0, ref, 0))
        ops.matmul_mask(dA1, dev(T, B1), bits, out=full[:, :y], transB=bool(tb), flags=ops.EPI_PADDED | ops.EPI_BITMASK)
        close(full[:, :y].cpu().numpy(), np.where(act > 0, ref1, 0))
        if not tb:
            bits2 = T.zeros(x, nw, dtype=T.int32, device="cuda")
            ops.matmul_relu_bits(dA1, dev(T, B1), bits2, out=full[:, :y], flags=ops.EPI_PADDED)
            a1 = full[:, :y].cpu().numpy()
            close(a1, np.maximum(ref1, 0))
            want2 = np.zeros((x, nw * 32), bool); want2[:, :y] = a1 > 0
            assert np.array_equal(np.unpackbits(bits2.cpu().numpy().view(np.uint8), axis=1, bitorder="little").astype(bool), want2)


@pytest.mark.parametrize("shape", [(20000, 256, 47, 47), (5000, 100, 64, 33), (4096, 128, 128, 100), (3000, 50, 7, 9)])
def test_matmul_ncat_vs_oracle(T, ops, liborc, shape):
    """C1 = A·B1, C2 = A·B2 with A streamed once (SAGE transform-first forward): two outputs with different row pitches."""
    x, z, y1, y2 = shape
    rng = np.random.default_rng(x + z + y1)
    A = rng.standard_normal((x, z), dtype=np.float32)
    B1 = rng.standard_normal((z, y1), dtype=np.float32)
    B2 = rng.standard_normal((z, y2), dtype=np.float32)
    r1, r2 = np.zeros((x, y1), np.float32), np.zeros((x, y2), np.float32)
    liborc.orc_gemm(x, y1, z, A.reshape(-1), B1.reshape(-1), r1.reshape(-1), 0, 0, 0)
    liborc.orc_gemm(x, y2, z, A.reshape(-1), B2.reshape(-1), r2.reshape(-1), 0, 0, 0)
    o1 = T.full((x, (y1 + 3) // 4 * 4), 7.0, device="cuda")
    o2 = T.empty(x, y2, device="cuda")
    ops.matmul_ncat(dev(T, A), dev(T, B1), dev(T, B2), out1=o1[:, :y1], out2=o2)
    close(o1[:, :y1].cpu().numpy(), r1)
    close(o2.cpu().numpy(), r2)
    assert bool((o1[:, y1:] == 7.0).all()), "padding columns of the first output must not be written"
    ops.matmul_ncat(dev(T, A), dev(T, B1), dev(T, B2), out1=o1[:, :y1], out2=o2, flags=ops.EPI_PADDED if y2 % 4 == 0 else 0)
    close(o1[:, :y1].cpu().numpy(), r1)
    close(o2.cpu().numpy(), r2)


@pytest.mark.parametrize("shape", [(50000, 100, 100, 256), (30011, 47, 128, 64), (4100, 16, 7, 33), (3000, 20, 30, 16)])
def test_wgrad_two_a_vs_fp64(T, ops, shape):
    """[dW_neigh; dW_self] = [ÂX | X]^T·dH with dH streamed once (sage_layer.cpp:37-47)."""
    n, x1, x2, y = shape
    rng = np.random.default_rng(n + x1)
    A1 = rng.standard_normal((n, x1), dtype=np.float32)
    A2 = rng.standard_normal((n, x2), dtype=np.float32)
    B = rng.standard_normal((n, y), dtype=np.float32)
    c1, c2 = ops.wgrad_two_a(dev(T, A1), dev(T, A2), dev(T, B))
    close(c1.cpu().numpy(), A1.astype(np.float64).T @ B.astype(np.float64), 5e-6)
    close(c2.cpu().numpy(), A2.astype(np.float64).T @ B.astype(np.float64), 5e-6)


@pytest.mark.parametrize("shape", [(50000, 256, 47, 47), (30011, 100, 64, 33), (4100, 200, 100, 128), (3000, 20, 30, 16)])
@pytest.mark.parametrize("padded", [False, True])
def test_wgrad_two_b_vs_fp64(T, ops, shape, padded):
    """[dW_neigh | dW_self] = H^T·[dY | dZ] with H streamed once; gradient operands in dense and 16-byte-padded row layouts."""
    n, x, y1, y2 = shape
    rng = np.random.default_rng(n + x + y1)
    A = rng.standard_normal((n, x), dtype=np.float32)
    pad = (lambda w: (w + 3) // 4 * 4) if padded else (lambda w: w)
    B1 = np.zeros((n, pad(y1)), np.float32); B1[:, :y1] = rng.standard_normal((n, y1), dtype=np.float32)
    B2 = np.zeros((n, pad(y2)), np.float32); B2[:, :y2] = rng.standard_normal((n, y2), dtype=np.float32)
    c1, c2 = ops.wgrad_two_b(dev(T, A), dev(T, B1)[:, :y1], dev(T, B2)[:, :y2])
    close(c1.cpu().numpy(), A.astype(np.float64).T @ B1[:, :y1].astype(np.float64), 5e-6)
    close(c2.cpu().numpy(), A.astype(np.float64).T @ B2[:, :y2].astype(np.float64), 5e-6)


def test_sigmoid_loss_and_f1_vs_reference_golden(T, ops):
    """Multi-label loss kernels against vectors from the live reference (tests/golden/sigmoid.npz): probabilities / losses / gradient within
    1e-6 (device expf / logf differ from glibc in the last bit), micro-F1 equal, through dense and pitched row layouts."""
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "sigmoid.npz"))
    x, y, m, b, e = z["x"], z["y"], z["m"], int(z["b"]), int(z["e"])
    nv, nc = x.shape
    sel = m.astype(bool); sel[:b] = False; sel[e:] = False
    for pitch in (nc, (nc + 3) // 4 * 4):
        logits = T.zeros(nv, pitch, device="cuda"); logits[:, :nc] = dev(T, x)
        probs = T.zeros(nv, pitch, device="cuda"); grad = T.zeros(nv, pitch, device="cuda"); losses = T.zeros(nv, device="cuda")
        dy, dm = dev(T, y), dev(T, m)
        ops.sigmoid_ce_forward(logits[:, :nc], dy, dm, b, e, probs[:, :nc], losses)
        ops.sigmoid_ce_backward(probs[:, :nc], dy, dm, b, e, grad[:, :nc])
        p = probs[:, :nc].cpu().numpy(); l = losses.cpu().numpy(); g = grad[:, :nc].cpu().numpy()
        np.testing.assert_allclose(p[sel], z["probs"][sel], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(l[sel], z["losses"][sel], rtol=2e-6)
        # p - y cancels where p ~ y: the bar is norm-wise (1e-6 of the largest gradient entry), as for every fp32 tensor in this suite
        np.testing.assert_allclose(g, z["grad"], rtol=1e-6, atol=1e-6 * float(np.abs(z["grad"]).max()))
        assert (g[~sel] == 0).all(), "rows outside the masked range keep their gradient"
        st = ops.masked_loss_mean(losses, dm, b, e).cpu().numpy()
        assert abs(st[0] - float(z["loss"])) <= 2e-6 * float(z["loss"]) and st[2] == sel.sum()
        f1 = float(ops.masked_f1_micro(probs[:, :nc], dy, dm, b, e).cpu()[0])
        assert abs(f1 - float(z["f1"])) < 1e-6


def test_dropout_mask_algebra_and_statistics(T, ops):
    """dropout_cpu / d_dropout_cpu (math_functions.cpp:417-440): out = in * mask * scale exactly, the mask is Bernoulli(1 - rate), redrawn
    per call, reproducible for the same (seed, call). The reference's generator is /dev/urandom-seeded, so only the algebra and the
    statistics can be pinned."""
    n, rate = 1 << 20, 0.3
    x = T.randn(n, device="cuda")
    out, mask = ops.dropout(x, rate, seed=7, call=0)
    scale = np.float32(1.0 / (1.0 - rate))
    want = (x.cpu().numpy() * mask.cpu().numpy().astype(np.float32)) * scale
    assert np.array_equal(out.cpu().numpy(), want)
    keep = float(mask.float().mean())
    assert abs(keep - (1 - rate)) < 3e-3, keep
    out2, mask2 = ops.dropout(x, rate, seed=7, call=0)
    assert T.equal(mask, mask2) and T.equal(out, out2)
    _, mask3 = ops.dropout(x, rate, seed=7, call=1)
    agree = float((mask3 == mask).float().mean())
    assert abs(agree - ((1 - rate) ** 2 + rate ** 2)) < 5e-3, agree  # independent redraw
    g = T.randn(n, device="cuda")
    dg = ops.d_dropout(g, mask, rate)
    assert np.array_equal(dg.cpu().numpy(), (g.cpu().numpy() * mask.cpu().numpy().astype(np.float32)) * scale)
