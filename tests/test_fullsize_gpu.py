"""Parity at BASELINE.json's full size (configs[1]: 2 449 029 vertices, ~62 M CSR edges) through size-independent properties: the
oracle cannot run the whole graph in seconds, so the kernels are checked against (a) closed forms that depend only on a row's degree,
(b) each other (different kernel instantiations must produce the same bits for the same column), (c) the oracle's arithmetic restated
in numpy on sampled rows — hub rows included — and (d) run-to-run determinism of the whole training step."""
import numpy as np
import pytest

from conftest import require_cuda

pytestmark = pytest.mark.gpu
NV, NNZ = 2_449_029, 62_000_000


@pytest.fixture(scope="module")
def big():
    require_cuda()
    import torch
    from graphaibench_b200 import build, datagen, ops
    build.build_all()
    rp, ci = datagen.rmat_csr_torch(NV, NNZ, seed=1, device="cuda")
    rp32, ci32 = rp.to(torch.int32), ci.to(torch.int32)
    g = ops.DeviceGraph(rp32, ci32, device_arrays=True)
    deg = (rp[1:] - rp[:-1]).cpu().numpy()
    return dict(T=torch, ops=ops, g=g, rowptr=rp.cpu().numpy(), colidx=ci.cpu().numpy(), deg=deg)


def seq_sum_f32(vals):
    """fp32 sequential sum, the reference's accumulation order (gcn_aggregator.cpp:56-73)."""
    return np.cumsum(vals.astype(np.float32), dtype=np.float32)[-1] if len(vals) else np.float32(0)


def test_mean_of_ones_is_the_degree_closed_form(big):
    """out_i = sum over deg_i edges of fl(1/deg_i) * 1, added in order: a function of the degree alone. Checked bit for bit on EVERY row
    (light rows, hub items and empty rows) against the fp32 running sum computed on the host once per distinct degree."""
    T, ops, g, deg = big["T"], big["ops"], big["g"], big["deg"]
    x = T.ones(NV, 8, device="cuda")
    out = ops.spmm_mean(g, x).cpu().numpy()
    assert (out == out[:, :1]).all(), "columns of an all-ones input must agree"
    dmax = int(deg.max())
    assert dmax > 8192, "the full-size graph must contain hub rows"
    want_by_deg = {}
    for d in np.unique(deg):
        if d == 0:
            want_by_deg[0] = np.float32(0)
            continue
        w = np.float32(1.0 / float(np.float32(d)))  # (float)(1.0 / (double)(float)deg), sage_aggregator.cpp:17
        want_by_deg[int(d)] = seq_sum_f32(np.full(int(d), w, np.float32))
    want = np.array([want_by_deg[int(d)] for d in deg], np.float32)
    assert np.array_equal(out[:, 0], want)


def test_width_instantiations_agree_bitwise(big):
    """The F = 100 kernel (32 lanes per row), the F = 47 kernel (16 lanes, padded rows) and the F = 8 kernel (4 lanes) must produce the
    same bits for the same column: per column the sum is sequential in edge order whatever the lane mapping."""
    T, ops, g = big["T"], big["ops"], big["g"]
    gen = T.Generator(device="cuda"); gen.manual_seed(5)
    x = T.randn(NV, 100, generator=gen, device="cuda")
    full = ops.spmm_mean(g, x)
    x48 = T.zeros(NV, 48, device="cuda"); x48[:, :47] = x[:, :47]
    o47 = ops.spmm_mean(g, x48[:, :47], out=T.empty(NV, 48, device="cuda")[:, :47])
    assert T.equal(o47, full[:, :47])
    o8 = ops.spmm_mean(g, x[:, 40:48].contiguous())
    assert T.equal(o8, full[:, 40:48])
    gcn = ops.spmm_gcn(g, x)
    gcn8 = ops.spmm_gcn(g, x[:, 40:48].contiguous())
    assert T.equal(gcn8, gcn[:, 40:48])
    tr = ops.spmm_mean(g, x, transposed=True)
    tr8 = ops.spmm_mean(g, x[:, 40:48].contiguous(), transposed=True)
    assert T.equal(tr8, tr[:, 40:48])


def test_sampled_rows_match_reference_arithmetic(big):
    """2 000 random rows + the 8 longest rows: acc = fadd(acc, fmul(w, x)) in CSR order, restated in numpy, bit-exact (GCN and both SAGE forms)."""
    T, ops, g = big["T"], big["ops"], big["g"]
    rowptr, colidx, deg = big["rowptr"], big["colidx"], big["deg"]
    gen = T.Generator(device="cuda"); gen.manual_seed(6)
    x = T.randn(NV, 12, generator=gen, device="cuda")
    xh = x.cpu().numpy()
    outs = {"gcn": ops.spmm_gcn(g, x).cpu().numpy(), "mean": ops.spmm_mean(g, x).cpu().numpy(),
            "meanT": ops.spmm_mean(g, x, transposed=True).cpu().numpy()}
    degf = deg.astype(np.float32)
    with np.errstate(divide="ignore"):
        ngcn = np.where(deg > 0, (1.0 / np.sqrt(degf).astype(np.float64)).astype(np.float32), np.float32(0))  # lgraph.cpp:29-32
        nmean = np.where(deg > 0, (1.0 / degf.astype(np.float64)).astype(np.float32), np.float32(0))
    rng = np.random.default_rng(7)
    rows = np.concatenate([rng.integers(0, NV, 2000), np.argsort(deg)[-8:]])
    for i in rows:
        cols = colidx[rowptr[i]:rowptr[i + 1]]
        xs = xh[cols]
        for name, w in (("gcn", (ngcn[i] * ngcn[cols]).astype(np.float32)), ("mean", np.full(len(cols), nmean[i], np.float32)),
                        ("meanT", nmean[cols])):
            prod = (w[:, None] * xs).astype(np.float32)
            want = np.cumsum(prod, axis=0, dtype=np.float32)[-1] if len(cols) else np.zeros(12, np.float32)
            assert np.array_equal(outs[name][i], want), (name, int(i), int(deg[i]))


def test_dense_transforms_at_full_height(big):
    """Tensor-core transforms over all 2.45 M rows: sampled rows against fp64, and a checksum of checksums (column sums of C against
    (column sums of A)·W in fp64) so that no tile can be skipped or written twice."""
    T, ops = big["T"], big["ops"]
    gen = T.Generator(device="cuda"); gen.manual_seed(8)
    a1, a2 = T.randn(NV, 100, generator=gen, device="cuda"), T.randn(NV, 100, generator=gen, device="cuda")
    w1, w2 = T.randn(100, 256, generator=gen, device="cuda") * 0.1, T.randn(100, 256, generator=gen, device="cuda") * 0.1
    c = ops.matmul_kcat(a1, w1, a2, w2)
    idx = T.randint(0, NV, (4000,), generator=gen, device="cuda")
    ref = a1[idx].double() @ w1.double() + a2[idx].double() @ w2.double()
    err = (c[idx].double() - ref).abs().max() / ref.abs().max()
    assert float(err) <= 1e-5, float(err)
    cs = c.double().sum(0)
    cs_ref = a1.double().sum(0) @ w1.double() + a2.double().sum(0) @ w2.double()
    scale = (a1.double().abs().sum(0) @ w1.double().abs() + a2.double().abs().sum(0) @ w2.double().abs())
    assert float(((cs - cs_ref).abs() / scale).max()) <= 1e-6
    # weight gradients: the reduction runs over all rows
    gmat = T.randn(NV, 47, generator=gen, device="cuda")
    d1, d2 = ops.wgrad_two_a(a1, a2, gmat)
    r1, r2 = a1.double().t() @ gmat.double(), a2.double().t() @ gmat.double()
    norm = float((a1.double().abs().t() @ gmat.double().abs()).max())
    assert float((d1.double() - r1).abs().max()) / norm <= 1e-6 and float((d2.double() - r2).abs().max()) / norm <= 1e-6


def test_training_step_is_deterministic_at_full_size(big):
    """Two models built from the same inputs produce identical losses, accuracies and weights bit for bit over three epochs (persistent
    kernels with dynamic work claims, hub items and multi-CTA reductions included: every reduction has a fixed order)."""
    T = big["T"]
    from graphaibench_b200 import datagen, model as gmodel
    gen = T.Generator(device="cuda"); gen.manual_seed(2)
    feats = T.randn(NV, 100, generator=gen, device="cuda").cpu().numpy()
    labels = T.randint(0, 47, (NV,), generator=gen, device="cuda").to(T.uint8).cpu().numpy()
    rp, ci = big["rowptr"].astype(np.uint32), big["colidx"].astype(np.uint32)
    runs = []
    for _ in range(2):
        m = gmodel.GnnModel("sage", rp, ci, feats, labels, datagen.split_ranges(NV), 256, 47, num_layers=2, lr=0.01)
        hist = [m.train_epoch() for _ in range(3)]
        runs.append((hist, m.get("W", 0), m.get("W_self", 1)))
    assert runs[0][0] == runs[1][0]
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    losses = [h[0] for h in runs[0][0]]
    assert all(np.isfinite(losses)), losses
