"""the reference's device-graph accessor surface (class GraphGPU, include/graph_gpu.h) over this library's device CSR
(include/gai_graph_gpu.cuh), exercised by the consumer the survey names — the vertex-parallel triangle kernel of src/triangle
(bs_warp_vertex.cuh), run on the whole graph and, as src/triangle/multigpu_induced.cu does, on every rank's induced subgraph of the 1D
partition (gai_partition1d_h, bit-exact vs the reference partitioner) over the rank's master rows."""
import ctypes as C

import numpy as np
import pytest

from conftest import require_cuda

pytestmark = pytest.mark.gpu


def wedge_closures(rp, ci, n):
    """sum_v sum_{u in N(v)} |N(v) ∩ N(u)| = sum((A·A) ∘ A), A the 0/1 adjacency (6 x triangles on a symmetric loop-free graph)."""
    import scipy.sparse as sp
    A = sp.csr_matrix((np.ones(len(ci), np.int64), ci.astype(np.int64), rp.astype(np.int64)), shape=(n, n))
    return int((A @ A).multiply(A).sum())


def count_rows(ops, g, begin, end):
    from graphaibench_b200._abi import check, lib
    out = C.c_uint64()
    check(lib().gai_triangle_count_rows(g.handle, begin, end, C.byref(out), None), "gai_triangle_count_rows")
    return out.value


@pytest.mark.parametrize("which", ["cora", "small"])
def test_triangle_consumer_on_whole_and_partitioned_graph(cora, small_graph, which):
    require_cuda()
    from graphaibench_b200 import build, ops
    build.build_all()
    rp64, ci = (cora["rowptr64"], cora["colidx"]) if which == "cora" else (small_graph["rowptr64"], small_graph["colidx"])
    n = len(rp64) - 1
    want = wedge_closures(rp64, ci, n)
    assert want > 0 and want % 6 == 0
    g = ops.DeviceGraph(rp64.astype(np.uint32), ci)
    from graphaibench_b200._abi import lib
    assert lib().gai_csr_max_degree(g.handle) == int(np.diff(rp64).max())
    assert count_rows(ops, g, 0, n) == want
    # ragged row ranges add up
    cuts = [0, n // 3, n // 3, n - 1, n]
    assert sum(count_rows(ops, g, a, b) for a, b in zip(cuts[:-1], cuts[1:])) == want
    # one induced subgraph per rank, counted over its masters only: the per-rank counts add up to the whole (multigpu_induced.cu)
    for world in (2, 4):
        total = 0
        for rank in range(world):
            part = ops.partition1d(rp64, ci, world, rank)
            if len(part["idx_map"]) == 0:
                continue
            sub = ops.DeviceGraph(part["rowptr"].astype(np.uint32), part["colidx"])
            total += count_rows(ops, sub, part["local_begin"], part["local_end"])
        assert total == want
