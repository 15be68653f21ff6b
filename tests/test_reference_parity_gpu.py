"""Reference-vs-GPU parity on the BASELINE.json shapes above cora (VERDICT r1 "weak" #1). The checker is the REFERENCE ITSELF
(oracle/_ref/libref_gnn.so = the reference's Model<L>, layers, aggregators and OpenBLAS sgemm compiled from its own sources; it travels
to the GPU box as a prebuilt binary), falling back to the C restatement only where that binary is missing.

  * configs[1] at FULL size (2 449 029 vertices, ~62 M CSR edges, SAGE 100 -> 256 -> 47): one training step on both sides, loss,
    all four weight gradients and 4 000 sampled activation / gradient rows.
  * configs[2]-shaped (GAT 602 -> 256 -> 256 + l2norm + dense -> 41) and configs[3]-shaped (GCN 128 -> 256 -> 256 -> 172, three layers)
    models on down-scaled R-MAT graphs that still contain hub rows (degree > the 1 024-edge hub threshold) and empty rows.
fp32 tolerances are norm-wise relative errors, written at each comparison."""
import os

import numpy as np
import pytest

from conftest import require_cuda

pytestmark = pytest.mark.gpu


def relerr(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30)


def close(a, ref, tol, what=""):
    err = relerr(a, ref)
    assert err <= tol, f"{what}: norm-wise relative error {err:.3e} > {tol}"


@pytest.fixture(scope="module")
def env():
    require_cuda()
    from graphaibench_b200 import build
    build.build_all()
    import oracle
    from graphaibench_b200 import datagen, model
    from oracle import model as om
    threads = os.cpu_count() or 1

    def checker(arch, rp, ci, feats, labels, split, hid, ncls, layers, lr):
        if oracle.have_ref():
            return oracle.RefModel(arch, rp, ci, feats, labels, split, hid, ncls, num_layers=layers, lr=lr, threads=threads), "reference"
        return om.OracleModel(arch, rp, ci, feats, labels, split, hid, ncls, num_layers=layers, lr=lr), "restatement"
    return dict(model=model, datagen=datagen, checker=checker, oracle=oracle)


def tensor(chk, kind, name, layer):
    if kind == "reference":
        return chk.get(name, layer)
    y = chk.layers[layer]
    return np.asarray(getattr(y, name)).ravel()


def test_c2_full_size_step_matches_reference(env):
    """BASELINE.json configs[1] at full size, the exact bench.py workload (same generator and seeds)."""
    import torch
    import bench
    w = bench.make_workload(1, "cuda")
    C2 = bench.C2
    nv = w["nv"]
    m = env["model"].GnnModel("sage", w["rowptr"], w["colidx"], w["feats"], w["labels"], w["split"], C2["hid"], C2["ncls"], num_layers=2, lr=C2["lr"])
    chk, kind = env["checker"]("sage", w["rowptr"], w["colidx"], w["feats"], w["labels"], w["split"], C2["hid"], C2["ncls"], 2, C2["lr"])
    l, a = m.forward(); lr_, ar_ = chk.forward()
    assert abs(l - lr_) <= 1e-5 * abs(lr_), (l, lr_)
    assert abs(a - ar_) <= 2e-6, (a, ar_)   # accuracy = correct / 1 224 514 rows: at most a couple of argmax ties may flip
    m.backward(); chk.backward()
    rows = np.random.default_rng(17).integers(0, nv, 4000)
    hubs = np.argsort(np.diff(w["rowptr"].astype(np.int64)))[-8:]
    rows = np.concatenate([rows, hubs])
    for k, width_in, width_out in ((0, C2["feat"], C2["hid"]), (1, C2["hid"], C2["ncls"])):
        close(m.get("W_grad", k), tensor(chk, kind, "W_grad", k), 2e-5, f"W_grad[{k}]")
        close(m.get("W_self_grad", k), tensor(chk, kind, "W_self_grad", k), 2e-5, f"W_self_grad[{k}]")
        if k > 0:
            close(m.get("feat_in", k).reshape(nv, width_in)[rows], tensor(chk, kind, "feat_in", k).reshape(nv, width_in)[rows], 1e-5, f"feat_in[{k}]")
        close(m.get("grad_in", k).reshape(nv, width_out)[rows], tensor(chk, kind, "grad_in", k).reshape(nv, width_out)[rows], 2e-5, f"grad_in[{k}]")
    m.update(); chk.update()
    l2, _ = m.train_epoch(); l2r, _ = chk.train_epoch()
    assert abs(l2 - l2r) <= 1e-4 * abs(l2r), (l2, l2r)   # second epoch: one Adam step apart from bit-identical initial weights


def _scaled_case(env, arch, nv, nnz, dims, layers, seed):
    dg = env["datagen"]
    rp64, ci = dg.rmat_csr(nv, nnz, seed=seed)
    rp = rp64.astype(np.uint32)
    deg = np.diff(rp64)
    assert deg.max() > 1024, "the down-scaled graph must keep hub rows (separate kernel path)"
    assert (deg == 0).any(), "and empty rows"
    F, hid, ncls = dims
    feats = dg.features(nv, F, seed=seed + 1)
    labels = np.random.default_rng(seed + 2).integers(0, ncls, nv).astype(np.uint8)
    split = dg.split_ranges(nv)
    m = env["model"].GnnModel(arch, rp, ci, feats, labels, split, hid, ncls, num_layers=layers, lr=0.01)
    chk, kind = env["checker"](arch, rp, ci, feats, labels, split, hid, ncls, layers, 0.01)
    return m, chk, kind, nv


def test_c3_shaped_gat_matches_reference(env):
    """configs[2] shape: 602 features, hidden 256, 2 GAT layers, l2norm + dense -> 41 classes, average degree ~100 with hub rows."""
    m, chk, kind, nv = _scaled_case(env, "gat", 16000, 1_600_000, (602, 256, 41), 2, seed=31)
    l, a = m.forward(); lr_, ar_ = chk.forward()
    assert abs(l - lr_) <= 1e-5 * abs(lr_), (l, lr_)
    assert abs(a - ar_) <= 1e-3
    m.backward(); chk.backward()
    for k in range(2):
        if k > 0:
            close(m.get("feat_in", k), tensor(chk, kind, "feat_in", k), 2e-5, f"feat_in[{k}]")
        close(m.get("grad_in", k), tensor(chk, kind, "grad_in", k), 5e-5, f"grad_in[{k}]")
        close(m.get("W_grad", k), tensor(chk, kind, "W_grad", k), 5e-5, f"W_grad[{k}]")
        if kind == "reference":
            close(m.get("alpha_lgrad", k), chk.get("alpha_lgrad", k), 1e-4, f"alpha_lgrad[{k}]")
            close(m.get("alpha_rgrad", k), chk.get("alpha_rgrad", k), 1e-4, f"alpha_rgrad[{k}]")
    if kind == "reference":
        close(m.get("dense_W_grad", 0), chk.get("dense_W_grad", 0), 2e-5, "dense_W_grad")


def test_c4_shaped_gcn3_matches_reference(env):
    """configs[3] shape: GCN 128 -> 256 -> 256 -> 172, three layers (aggregate-first, aggregate-first with the sign-bit d_relu epilogue,
    transform-first with 172-class rows)."""
    m, chk, kind, nv = _scaled_case(env, "gcn", 40000, 1_200_000, (128, 256, 172), 3, seed=41)
    l, a = m.forward(); lr_, ar_ = chk.forward()
    assert abs(l - lr_) <= 1e-5 * abs(lr_), (l, lr_)
    assert abs(a - ar_) <= 1e-3
    m.backward(); chk.backward()
    for k in range(3):
        if k > 0:
            close(m.get("feat_in", k), tensor(chk, kind, "feat_in", k), 1e-5, f"feat_in[{k}]")
        close(m.get("grad_in", k), tensor(chk, kind, "grad_in", k), 2e-5, f"grad_in[{k}]")
        close(m.get("W_grad", k), tensor(chk, kind, "W_grad", k), 2e-5, f"W_grad[{k}]")
    m.update(); chk.update()
    for ep in range(2):
        l, _ = m.train_epoch(); lr_, _ = chk.train_epoch()
        assert abs(l - lr_) <= 1e-4 * abs(lr_), (ep, l, lr_)
