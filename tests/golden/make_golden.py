"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libref_gnn.so + libref_part.so, built from
/root/reference's own sources by oracle/build_ref.sh). Run in the build container only:

    python tests/golden/make_golden.py

The fixtures are what pins the C restatement (oracle/gnn_oracle.c) and the CUDA path on boxes where /root/reference
does not exist.  Bit-exact quantities are stored as arrays or SHA-256 digests; tolerance quantities as arrays.
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import model as om  # noqa: E402
from graphaibench_b200 import datagen  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF_INPUTS = "/root/reference/inputs"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_graph(seed=11, n=1500, nnz=24000, hub=1200):
    """Power-law graph + an explicit hub row (> HUB_DEGREE=1024 neighbours) + isolated vertices."""
    rp, ci = datagen.rmat_csr(n, nnz, seed=seed)
    import scipy.sparse as sp
    A = sp.csr_matrix((np.ones(len(ci), np.int8), ci.astype(np.int64), rp), shape=(n, n)).tolil()
    rng = np.random.default_rng(seed + 1)
    nb = rng.choice(np.arange(1, n), hub, replace=False)
    for j in nb:
        A[0, j] = 1; A[j, 0] = 1
    for v in (7, 8, n - 1):  # isolated vertices (degree 0 rows)
        A[v, :] = 0; A[:, v] = 0
    A = A.tocsr(); A.eliminate_zeros(); A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.uint32)


def main():
    assert oracle.have_ref(), "run oracle/build_ref.sh first"
    L = oracle.libref()
    L.ref_set_threads(1)

    # ---- cora dataset, stored sparsely (features are 1.27% dense) ------------------------------------------
    d = om.read_dataset(os.path.join(REF_INPUTS, "cora"))
    feats = d["feats"]
    nzr, nzc = np.nonzero(feats)
    np.savez_compressed(os.path.join(OUT, "cora.npz"), rowptr64=d["rowptr64"], colidx=d["colidx"], labels=d["labels"], split=d["split"],
                        feat_shape=np.array(feats.shape), feat_rows=nzr.astype(np.uint16), feat_cols=nzc.astype(np.uint16),
                        feat_vals=feats[nzr, nzc], ncls=d["ncls"], max_degree=d["max_degree"])

    gold = {}
    # ---- init_glorot (math_functions.cpp:11-19) ------------------------------------------------------------
    for (dx, dy, seed) in ((1433, 16, 1), (16, 7, 1), (16, 1, 2), (16, 1, 3), (100, 256, 2)):
        w = np.zeros(dx * dy, np.float32); L.ref_init_glorot(dx, dy, w, seed)
        gold[f"glorot_{dx}_{dy}_{seed}"] = w if dx * dy <= 1024 else w[:64].copy()
        gold[f"glorot_{dx}_{dy}_{seed}_sha"] = sha(w)

    # ---- graph prep + aggregators on a small power-law graph with a hub row and isolated vertices -----------
    rp64, ci = small_graph()
    rp = rp64.astype(np.uint32)
    n = len(rp) - 1
    gold["sg_rowptr64"], gold["sg_colidx"] = rp64, ci
    g_raw = L.ref_graph_new(n, len(ci), rp, ci)
    g_loop = L.ref_graph_new(n, len(ci), rp, ci)
    L.ref_graph_add_selfloop(g_loop)
    L.ref_graph_compute_vertex_data(g_loop)
    L.ref_graph_compute_vertex_data(g_raw)
    ne2 = L.ref_graph_ne(g_loop)
    rp2, ci2, vd2 = np.zeros(n + 1, np.uint32), np.zeros(ne2, np.uint32), np.zeros(n, np.float32)
    L.ref_graph_export(g_loop, rp2, ci2, vd2.ctypes.data_as(C.c_void_p), None)
    gold["sg_loop_rowptr"], gold["sg_loop_colidx_sha"], gold["sg_loop_vdata"] = rp2, sha(ci2), vd2
    rng = np.random.default_rng(5)
    for F in (7, 16, 47, 100, 256):
        x = rng.standard_normal((n, F), dtype=np.float32)
        gold[f"sg_x_{F}_sha"] = sha(x)  # inputs are regenerated from the seed in the tests; digest guards the RNG
        out = np.zeros((n, F), np.float32)
        L.ref_gcn_aggregate(g_loop, F, x.reshape(-1), out.reshape(-1)); gold[f"sg_gcn_{F}_sha"] = sha(out)
        if F == 16: gold["sg_gcn_16"] = out.copy()
        L.ref_sage_aggregate(g_raw, F, x.reshape(-1), out.reshape(-1), 0); gold[f"sg_mean_{F}_sha"] = sha(out)
        L.ref_sage_aggregate(g_raw, F, x.reshape(-1), out.reshape(-1), 1); gold[f"sg_meanT_{F}_sha"] = sha(out)
    # GAT aggregator forward/backward (tolerance): F = 16
    F = 16
    z = rng.standard_normal((n, F), dtype=np.float32) * 0.5
    gin = rng.standard_normal((n, F), dtype=np.float32)
    al = rng.standard_normal(F, dtype=np.float32) * 0.3
    ar = rng.standard_normal(F, dtype=np.float32) * 0.3
    out = np.zeros((n, F), np.float32); ns = np.zeros(ne2, np.float32); gout = np.zeros((n, F), np.float32)
    dal, dar = np.zeros(F, np.float32), np.zeros(F, np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    L.ref_gat_aggregate(g_loop, F, al, ar, z.reshape(-1), out.reshape(-1), vp(ns), vp(gin), vp(gout), vp(dal), vp(dar))
    gold.update(gat_z=z, gat_gin=gin, gat_al=al, gat_ar=ar, gat_out=out, gat_norm_scores=ns, gat_gout=gout, gat_dal=dal, gat_dar=dar)
    # transpose of edge values on the symmetric pattern
    vals = rng.standard_normal(ne2, dtype=np.float32)
    vt = np.zeros(ne2, np.float32)
    L.ref_symmetric_csr_transpose(n, ne2, rp2, ci2, vals, vt)
    gold["sg_vals_sha"], gold["sg_valsT_sha"] = sha(vals), sha(vt)

    # ---- loss / adam ----------------------------------------------------------------------------------------
    ncls, nv = 7, 64
    logits = rng.standard_normal((nv, ncls), dtype=np.float32) * 3
    labs = rng.integers(0, ncls, nv, dtype=np.uint8)
    masks = np.zeros(nv, np.uint8); masks[5:40] = 1
    probs = np.zeros((nv, ncls), np.float32); grad = np.zeros((nv, ncls), np.float32); acc = C.c_float()
    loss = L.ref_softmax_loss(nv, ncls, logits.reshape(-1), labs, vp(masks), 5, 40, 35, vp(probs), vp(grad), C.byref(acc))
    gold.update(loss_logits=logits, loss_labels=labs, loss_masks=masks, loss_probs=probs, loss_grad=grad,
                loss_value=np.float32(loss), loss_acc=np.float32(acc.value))
    nW, steps = 257, 3
    W = rng.standard_normal(nW, dtype=np.float32); grads = rng.standard_normal((steps, nW), dtype=np.float32) * 0.1
    W_after = W.copy(); L.ref_adam_steps(nW, 0.02, steps, grads.reshape(-1), W_after)
    gold.update(adam_W=W, adam_grads=grads, adam_W_after=W_after)

    # ---- end-to-end training on cora: per-epoch loss/acc, tensors after epoch 0, final accuracies -------------
    for arch, epochs in (("gcn", 200), ("sage", 100), ("gat", 100)):
        m = oracle.RefModel(arch, d["rowptr"], d["colidx"], feats, d["labels"], d["split"], 16, d["ncls"], threads=1)
        losses, accs = [], []
        for ep in range(epochs):
            if ep == 0:
                l, a = m.forward()
                gold[f"cora_{arch}_logits0"] = m.get("logits")
                m.backward()
                gold[f"cora_{arch}_Wgrad0_l0_sha"] = sha(m.get("W_grad", 0))
                gold[f"cora_{arch}_Wgrad0_l1"] = m.get("W_grad", 1)
                gold[f"cora_{arch}_Wgrad0_l0_sample"] = m.get("W_grad", 0)[::97].copy()
                gold[f"cora_{arch}_gradin0_l0_sample"] = m.get("grad_in", 0)[::101].copy()
                if arch == "gat":
                    gold["cora_gat_alpha_lgrad0_l0"] = m.get("alpha_lgrad", 0)
                    gold["cora_gat_alpha_rgrad0_l0"] = m.get("alpha_rgrad", 0)
                m.update()
                gold[f"cora_{arch}_W1_l0_sample"] = m.get("W", 0)[::97].copy()
            else:
                l, a = m.train_epoch()
            losses.append(l); accs.append(a)
        gold[f"cora_{arch}_losses"] = np.array(losses, np.float32)
        gold[f"cora_{arch}_accs"] = np.array(accs, np.float32)
        gold[f"cora_{arch}_test_acc"] = np.float32(m.evaluate("test"))
        gold[f"cora_{arch}_val_acc"] = np.float32(m.evaluate("val"))
        print(arch, "final loss", losses[-1], "test", gold[f"cora_{arch}_test_acc"], "val", gold[f"cora_{arch}_val_acc"])

    # ---- partitioner (integer goldens) ----------------------------------------------------------------------
    if os.path.exists(oracle._LIBREFPART_PATH):
        P = C.CDLL(oracle._LIBREFPART_PATH)
        for name, (prp, pci) in (("cora", (d["rowptr64"], d["colidx"])), ("sg", (rp64, ci))):
            for nparts in (2, 4):
                for part in range(nparts):
                    res = oracle.ref_partition1d(prp, pci, nparts, part)
                    key = f"part_{name}_{nparts}_{part}"
                    gold[key + "_lb_le_m_ne"] = np.array([res["local_begin"], res["local_end"], len(res["idx_map"]), len(res["colidx"])], np.int64)
                    gold[key + "_idx_sha"], gold[key + "_rowptr_sha"], gold[key + "_colidx_sha"] = sha(res["idx_map"]), sha(res["rowptr"]), sha(res["colidx"])
    np.savez_compressed(os.path.join(OUT, "golden.npz"), **gold)
    for f in ("cora.npz", "golden.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
