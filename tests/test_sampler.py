"""subgraph sampling. The reference's Sampler (src/gnn/sampler.cpp, compiled into oracle/_ref/libref_gnn.so, which travels
to the GPU box) against host/gai_sampler.cpp: the frontier walk (select_vertices) must pick the SAME vertex set from the same seed — it is
restated operation for operation over the same libc rand_r stream — on CPU; the induced, re-indexed subgraph built on the device
(csrc/convert.cu: gai_induced_subgraph) must be the reference's CSR bit for bit, on the GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import require_cuda


def _graph(nv, nnz, seed):
    from graphaibench_b200 import datagen
    rp64, ci = datagen.rmat_csr(nv, nnz, seed=seed)
    masks = np.zeros(nv, np.uint8)
    masks[: nv // 2] = 1
    return rp64.astype(np.uint32), ci, masks


def _ref(rp, ci, masks, n, seed):
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libref_gnn.so not built")
    L = oracle.libref()
    L.ref_sampler_run.restype = C.c_int64
    L.ref_sampler_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    g = L.ref_graph_new(len(rp) - 1, len(ci), rp, ci)
    sizes = np.zeros(2, np.int64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    L.ref_sampler_run(g, vp(masks), int(masks.sum()), n, seed, vp(sizes), None, None, None)
    st, srp, sci = np.zeros(sizes[0], np.uint32), np.zeros(sizes[0] + 1, np.uint32), np.zeros(max(sizes[1], 1), np.uint32)
    g = L.ref_graph_new(len(rp) - 1, len(ci), rp, ci)
    L.ref_sampler_run(g, vp(masks), int(masks.sum()), n, seed, vp(sizes), vp(st), vp(srp), vp(sci))
    return st, srp, sci[: sizes[1]]


def _ours(rp, ci, masks, n, seed, with_subgraph):
    from graphaibench_b200 import model
    L = model.hostlib()
    L.gai_sampler_run.restype = C.c_int64
    L.gai_sampler_run.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    sizes = np.zeros(2, np.int64)
    L.gai_sampler_run(len(rp) - 1, vp(rp), vp(ci), vp(masks), int(masks.sum()), n, seed, int(with_subgraph), vp(sizes), None, None, None)
    st, srp, sci = np.zeros(sizes[0], np.uint32), np.zeros(sizes[0] + 1, np.uint32), np.zeros(max(sizes[1], 1), np.uint32)
    L.gai_sampler_run(len(rp) - 1, vp(rp), vp(ci), vp(masks), int(masks.sum()), n, seed, int(with_subgraph), vp(sizes), vp(st), vp(srp), vp(sci))
    return st, srp, sci[: sizes[1]]


# frontier smaller than, equal to and larger than the subgraph; dashboards that overflow their first reservation (compaction path)
CASES = [(3000, 40000, 100, 0), (20000, 400000, 3000, 1), (20000, 400000, 5000, 2), (60000, 2_000_000, 9000, 3), (30000, 300000, 12000, 7)]


@pytest.mark.parametrize("nv,nnz,n,seed", CASES)
def test_select_vertices_matches_reference_walk(nv, nnz, n, seed):
    rp, ci, masks = _graph(nv, nnz, 100 + seed)
    want, _, _ = _ref(rp, ci, masks, n, seed)
    got, _, _ = _ours(rp, ci, masks, n, seed, with_subgraph=False)
    assert len(want) > 0 and np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("nv,nnz,n,seed", CASES[1:4])
def test_device_induced_subgraph_matches_reference(nv, nnz, n, seed):
    require_cuda()
    rp, ci, masks = _graph(nv, nnz, 100 + seed)
    want = _ref(rp, ci, masks, n, seed)
    got = _ours(rp, ci, masks, n, seed, with_subgraph=True)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert len(want[2]) > 0
