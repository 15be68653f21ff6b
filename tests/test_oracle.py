"""CPU suite, part 1: the oracle (oracle/gnn_oracle.c + oracle/model.py) against the golden vectors generated from the
reference itself, and against the reference build directly when oracle/_ref is present."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import model as om
from conftest import sha


def test_glorot_matches_reference(golden):
    for (dx, dy, seed) in ((1433, 16, 1), (16, 7, 1), (16, 1, 2), (16, 1, 3), (100, 256, 2)):
        w = om.glorot(dx, dy, seed)
        assert sha(w) == str(golden[f"glorot_{dx}_{dy}_{seed}_sha"])


def test_selfloop_and_norm_bit_exact(golden, small_graph):
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"])
    g.add_selfloop()
    g.compute_vertex_data()
    assert np.array_equal(g.rowptr, golden["sg_loop_rowptr"])
    assert sha(g.colidx) == str(golden["sg_loop_colidx_sha"])
    assert np.array_equal(g.vdata, golden["sg_loop_vdata"])


@pytest.mark.parametrize("F", [7, 16, 47, 100, 256])
def test_aggregators_bit_exact(golden, small_graph, liborc, F):
    n, x = small_graph["n"], small_graph["x"][F]
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"])
    graw = om.Graph(small_graph["rowptr"], small_graph["colidx"])
    g.add_selfloop(); g.compute_vertex_data()
    out = np.zeros((n, F), np.float32)
    liborc.orc_spmm_gcn(n, g.rowptr, g.colidx, g.vdata, F, x.reshape(-1), out.reshape(-1))
    assert sha(out) == str(golden[f"sg_gcn_{F}_sha"])
    liborc.orc_spmm_mean(n, graw.rowptr, graw.colidx, F, x.reshape(-1), out.reshape(-1), 0)
    assert sha(out) == str(golden[f"sg_mean_{F}_sha"])
    liborc.orc_spmm_mean(n, graw.rowptr, graw.colidx, F, x.reshape(-1), out.reshape(-1), 1)
    assert sha(out) == str(golden[f"sg_meanT_{F}_sha"])


def test_symmetric_transpose_bit_exact(golden, small_graph, liborc):
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.add_selfloop()
    rng = np.random.default_rng(5)
    for F in (7, 16, 47, 100, 256):
        rng.standard_normal((g.nv, F), dtype=np.float32)
    # values are regenerated exactly as make_golden.py does (same generator order)
    z = rng.standard_normal((g.nv, 16), dtype=np.float32); gin = rng.standard_normal((g.nv, 16), dtype=np.float32)
    al = rng.standard_normal(16, dtype=np.float32); ar = rng.standard_normal(16, dtype=np.float32)
    vals = rng.standard_normal(g.ne, dtype=np.float32)
    assert sha(vals) == str(golden["sg_vals_sha"])
    out = np.zeros(g.ne, np.float32)
    bad = liborc.orc_symmetric_transpose(g.nv, g.rowptr, g.colidx, vals.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), None)
    assert bad == 0 and sha(out) == str(golden["sg_valsT_sha"])
    perm = g.transpose_perm()
    out2 = np.zeros_like(out); out2[perm] = vals
    assert np.array_equal(out, out2)


def test_gat_aggregator_matches_reference(golden, small_graph, liborc):
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.add_selfloop()
    n, F = g.nv, 16
    z, gin, al, ar = golden["gat_z"], golden["gat_gin"], golden["gat_al"], golden["gat_ar"]
    ts, sc, ns = (np.zeros(g.ne, np.float32) for _ in range(3))
    out = np.zeros((n, F), np.float32)
    liborc.orc_gat_forward(n, g.rowptr, g.colidx, F, al, ar, 0.2, z.reshape(-1), ts, sc, ns, out.reshape(-1))
    assert np.array_equal(ns, golden["gat_norm_scores"])  # same libm on the same machine family: exact
    assert np.array_equal(out, golden["gat_out"])
    nsg, dal, dar = np.zeros(g.ne, np.float32), np.zeros(F, np.float32), np.zeros(F, np.float32)
    gout = np.zeros((n, F), np.float32)
    liborc.orc_gat_backward(n, g.rowptr, g.colidx, F, 0.2, z.reshape(-1), gin.reshape(-1), ts, ns, sc, nsg, dal, dar, gout.reshape(-1), 0)
    assert np.array_equal(gout, golden["gat_gout"])
    np.testing.assert_allclose(dal, golden["gat_dal"], rtol=0, atol=0)
    np.testing.assert_allclose(dar, golden["gat_dar"], rtol=0, atol=0)
    # the closed-form d_softmax branch (AVX512 build of the reference) agrees to rounding
    liborc.orc_gat_backward(n, g.rowptr, g.colidx, F, 0.2, z.reshape(-1), gin.reshape(-1), ts, ns, sc, nsg, dal, dar, gout.reshape(-1), 1)
    np.testing.assert_allclose(dal, golden["gat_dal"], rtol=2e-5, atol=1e-6)


def test_multi_head_attention_with_one_head_is_the_reference_path(golden, small_graph, liborc):
    """The multi-head extension (BASELINE.json configs[2] names 8 heads; the reference has one) is defined so that heads == 1 reproduces
    the pinned single-head restatement bit for bit, forward and backward; with H > 1 every head must equal a single-head run on its own
    block of columns (same graph, that block's attention vectors)."""
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.add_selfloop()
    n, F = g.nv, 16
    z, gin, al, ar = golden["gat_z"], golden["gat_gin"], golden["gat_al"], golden["gat_ar"]

    def run(heads, fn_fwd, fn_bwd, zz, gg, aal, aar, width):
        ts, sc, ns, nsg = (np.zeros(g.ne * (heads or 1), np.float32) for _ in range(4))
        out, gout = np.zeros((n, width), np.float32), np.zeros((n, width), np.float32)
        dal, dar = np.zeros(width, np.float32), np.zeros(width, np.float32)
        zz, gg = np.ascontiguousarray(zz), np.ascontiguousarray(gg)
        if heads is None:
            liborc.orc_gat_forward(n, g.rowptr, g.colidx, width, aal, aar, 0.2, zz.reshape(-1), ts, sc, ns, out.reshape(-1))
            liborc.orc_gat_backward(n, g.rowptr, g.colidx, width, 0.2, zz.reshape(-1), gg.reshape(-1), ts, ns, sc, nsg, dal, dar, gout.reshape(-1), 0)
        else:
            liborc.orc_gat_forward_heads(n, g.rowptr, g.colidx, width, heads, aal, aar, 0.2, zz.reshape(-1), ts, sc, ns, out.reshape(-1))
            liborc.orc_gat_backward_heads(n, g.rowptr, g.colidx, width, heads, 0.2, zz.reshape(-1), gg.reshape(-1), ts, ns, sc, nsg, dal, dar,
                                          gout.reshape(-1), 0)
        return ns, out, gout, dal, dar

    ref = run(None, None, None, z, gin, al, ar, F)
    one = run(1, None, None, z, gin, al, ar, F)
    for a, b in zip(ref, one):
        assert np.array_equal(a, b)
    assert np.array_equal(ref[1], golden["gat_out"]) and np.array_equal(ref[2], golden["gat_gout"])
    H, D = 4, F // 4
    multi = run(H, None, None, z, gin, al, ar, F)
    for h in range(H):
        blk = slice(h * D, (h + 1) * D)
        single = run(None, None, None, z[:, blk], gin[:, blk], np.ascontiguousarray(al[blk]), np.ascontiguousarray(ar[blk]), D)
        assert np.array_equal(multi[0].reshape(-1, H)[:, h], single[0])          # per-head softmax
        assert np.array_equal(multi[1][:, blk], single[1]) and np.array_equal(multi[2][:, blk], single[2])
        assert np.array_equal(multi[3][blk], single[3]) and np.array_equal(multi[4][blk], single[4])


def test_loss_and_adam_bit_exact(golden, liborc):
    logits, labs, masks = golden["loss_logits"], golden["loss_labels"], golden["loss_masks"]
    nv, ncls = logits.shape
    probs, losses, grad = np.zeros((nv, ncls), np.float32), np.zeros(nv, np.float32), np.zeros((nv, ncls), np.float32)
    acc = C.c_float()
    loss = liborc.orc_softmax_loss(ncls, logits.reshape(-1), labs, masks.ctypes.data_as(C.c_void_p), 5, 40, probs.reshape(-1), losses,
                                   grad.ctypes.data_as(C.c_void_p), C.byref(acc))
    assert np.array_equal(probs, golden["loss_probs"]) and np.array_equal(grad, golden["loss_grad"])
    assert np.float32(loss) == golden["loss_value"] and np.float32(acc.value) == golden["loss_acc"]
    W = golden["adam_W"].copy()
    opt = om.Adam(0.02)
    for s in range(golden["adam_grads"].shape[0]):
        opt.update(golden["adam_grads"][s], W)
    assert np.array_equal(W, golden["adam_W_after"])


@pytest.mark.parametrize("arch,epochs", [("gcn", 200), ("sage", 100), ("gat", 100)])
def test_cora_training_matches_reference(golden, cora, arch, epochs):
    """Whole-model restatement: per-epoch loss trajectory and final accuracy equal the reference's (SURVEY.md §8c:
    0.795 / 0.784 / 0.771). Dense transforms differ from OpenBLAS in summation order, hence a tolerance on losses."""
    m = om.OracleModel(arch, cora["rowptr"], cora["colidx"], cora["feats"], cora["labels"], cora["split"], 16, cora["ncls"])
    losses = []
    for ep in range(epochs):
        if ep == 0:
            l, a = m.forward()
            ref_logits = golden[f"cora_{arch}_logits0"]
            assert np.abs(m.logits - ref_logits).max() <= 1e-5 * np.abs(ref_logits).max()
            m.backward()
            g1 = golden[f"cora_{arch}_Wgrad0_l1"]
            assert np.abs(m.layers[1].W_grad - g1).max() <= 1e-5 * np.abs(g1).max()
            m.update()
        else:
            l, a = m.train_epoch()
        losses.append(l)
    ref = golden[f"cora_{arch}_losses"]
    np.testing.assert_allclose(np.array(losses[:20], np.float32), ref[:20], rtol=1e-4)
    assert abs(m.evaluate("test") - float(golden[f"cora_{arch}_test_acc"])) < 1e-6
    assert abs(m.evaluate("val") - float(golden[f"cora_{arch}_val_acc"])) < 1e-6


def test_partition_restatement_bit_exact(golden, cora, small_graph):
    for name, (rp, ci) in (("cora", (cora["rowptr64"], cora["colidx"])), ("sg", (small_graph["rowptr64"], small_graph["colidx"]))):
        for nparts in (2, 4):
            for part in range(nparts):
                r = oracle.orc_partition1d(rp, ci, nparts, part)
                key = f"part_{name}_{nparts}_{part}"
                assert list(golden[key + "_lb_le_m_ne"]) == [r["local_begin"], r["local_end"], len(r["idx_map"]), len(r["colidx"])]
                assert sha(r["idx_map"]) == str(golden[key + "_idx_sha"])
                assert sha(r["rowptr"]) == str(golden[key + "_rowptr_sha"])
                assert sha(r["colidx"]) == str(golden[key + "_colidx_sha"])


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference in the build container)")
def test_restatement_against_live_reference(small_graph):
    """Direct check against the reference binary on fresh random inputs (beyond the committed goldens)."""
    L = oracle.libref(); L.ref_set_threads(2)
    rng = np.random.default_rng(99)
    n = small_graph["n"]
    g = om.Graph(small_graph["rowptr"], small_graph["colidx"]); g.add_selfloop(); g.compute_vertex_data()
    rg = L.ref_graph_new(n, len(small_graph["colidx"]), small_graph["rowptr"], small_graph["colidx"])
    L.ref_graph_add_selfloop(rg); L.ref_graph_compute_vertex_data(rg)
    for F in (3, 33, 130):
        x = rng.standard_normal((n, F), dtype=np.float32)
        a, b = np.zeros((n, F), np.float32), np.zeros((n, F), np.float32)
        L.ref_gcn_aggregate(rg, F, x.reshape(-1), a.reshape(-1))
        oracle.liborc().orc_spmm_gcn(n, g.rowptr, g.colidx, g.vdata, F, x.reshape(-1), b.reshape(-1))
        assert np.array_equal(a, b)
    A = rng.standard_normal((300, 70), dtype=np.float32); B = rng.standard_normal((70, 40), dtype=np.float32)
    c1, c2 = np.zeros((300, 40), np.float32), np.zeros((300, 40), np.float32)
    L.ref_matmul(300, 40, 70, A.reshape(-1), B.reshape(-1), c1.reshape(-1), 0, 0, 0)
    oracle.liborc().orc_gemm(300, 40, 70, A.reshape(-1), B.reshape(-1), c2.reshape(-1), 0, 0, 0)
    assert np.abs(c1 - c2).max() <= 1e-5 * np.abs(c1).max()


def test_sigmoid_loss_restatement_matches_reference_golden(liborc):
    """orc_sigmoid_loss (sigmoid_loss_layer.cpp:4-55, math_functions.cpp:517-521,553-559,580-623) against vectors generated from the live
    reference (tests/golden/make_golden_sigmoid.py): probabilities, per-row losses, gradient, mean loss and micro-F1, bit for bit."""
    import ctypes as C
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "sigmoid.npz"))
    x, y, m, b, e = z["x"], z["y"], z["m"], int(z["b"]), int(z["e"])
    nv, nc = x.shape
    probs = np.zeros((nv, nc), np.float32); losses = np.zeros(nv, np.float32); grad = np.zeros((nv, nc), np.float32); f1 = C.c_float()
    loss = liborc.orc_sigmoid_loss(nc, x.reshape(-1), y.reshape(-1), m.ctypes.data_as(C.c_void_p), b, e, probs.reshape(-1), losses,
                                   grad.ctypes.data_as(C.c_void_p), C.byref(f1))
    sel = m.astype(bool); sel[:b] = False; sel[e:] = False
    assert np.array_equal(probs[sel], z["probs"][sel]) and np.array_equal(losses[sel], z["losses"][sel]) and np.array_equal(grad, z["grad"])
    # the reference adds the per-row losses under an OpenMP reduction (sigmoid_loss_layer.cpp:38-47): order, hence the last bit, is free
    assert abs(np.float32(loss) - z["loss"]) <= 1e-6 * z["loss"] and np.float32(f1.value) == z["f1"]


@pytest.mark.parametrize("arch", ["gcn", "sage"])
def test_sigmoid_training_restatement_against_live_reference(cora, arch):
    """Whole-model restatement with the multi-label loss (argv[4] == "sigmoid": multi-hot labels, sigmoid_loss_layer, micro-F1 as
    accuracy; net.cpp:20,447-451,495-497,569-572) against the reference itself (oracle/_ref/libref_gnn.so, built from /root/reference):
    first-step tensors and 20 epochs of loss / F1. Skipped where the reference build is absent (the GPU box)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    args = (arch, cora["rowptr"], cora["colidx"], cora["feats"], cora["labels"], cora["split"], 16, cora["ncls"])
    m = om.OracleModel(*args, sigmoid=True)
    r = oracle.RefModel(*args, sigmoid=True)
    (l, a), (lr, ar) = m.forward(), r.forward()
    assert abs(l - lr) <= 1e-5 * abs(lr) and abs(a - ar) < 1e-6
    ref_logits = r.get("logits")
    assert np.abs(m.logits - ref_logits).max() <= 1e-5 * np.abs(ref_logits).max()
    m.backward(); r.backward()
    for k in (0, 1):
        g = r.get("W_grad", k)
        assert np.abs(m.layers[k].W_grad - g).max() <= 1e-5 * np.abs(g).max()
    m.update(); r.update()
    for ep in range(20):
        (l, a), (lr, ar) = m.train_epoch(), r.train_epoch()
        assert abs(l - lr) <= 2e-4 * abs(lr), (ep, l, lr)
        assert abs(a - ar) <= 0.02, (ep, a, ar)  # F1 thresholds predictions at 0.5: a last-bit difference can flip a handful
    assert abs(m.evaluate("test") - r.evaluate("test")) <= 0.01
