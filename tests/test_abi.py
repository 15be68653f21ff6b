"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/gai_b200.h declares, and its host-side
(integer) entry points are bit-exact against the golden vectors. No compute entry point is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, sha


@pytest.fixture(scope="module")
def abi():
    from graphaibench_b200 import build, _abi
    build.build_cuda()
    return _abi


def test_header_symbols_exported(abi):
    hdr = open(os.path.join(ROOT, "include", "gai_b200.h")).read()
    declared = set(re.findall(r"\b(gai_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = abi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in gai_b200.h but not exported by libgai_b200.so"
    assert declared == set(abi.EXPORTED), (declared ^ set(abi.EXPORTED))
    assert L.gai_version() >= 100


def test_library_is_sm100a_and_has_no_cpu_fallback(abi):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    # with no GPU present compute entry points fail loudly instead of silently computing on the host
    import torch
    if not torch.cuda.is_available():
        n = C.c_int(-1)
        rc = abi.lib().gai_device_count(C.byref(n))
        assert rc != 0 or n.value == 0
        p = C.c_void_p()
        assert abi.lib().gai_malloc(C.byref(p), 1024) != 0
        assert abi.lib().gai_last_error()


def test_add_selfloop_h_bit_exact(abi, golden, small_graph, cora):
    from graphaibench_b200 import ops
    rp, ci = ops.add_selfloop(small_graph["rowptr"], small_graph["colidx"])
    assert np.array_equal(rp, golden["sg_loop_rowptr"]) and sha(ci) == str(golden["sg_loop_colidx_sha"])
    # empty graph and single isolated vertex
    rp, ci = ops.add_selfloop(np.zeros(1, np.uint32), np.zeros(0, np.uint32))
    assert list(rp) == [0] and len(ci) == 0
    rp, ci = ops.add_selfloop(np.zeros(2, np.uint32), np.zeros(0, np.uint32))
    assert list(rp) == [0, 1] and list(ci) == [0]
    # against the oracle on cora
    from oracle import model as om
    g = om.Graph(cora["rowptr"], cora["colidx"]); g.add_selfloop()
    rp, ci = ops.add_selfloop(cora["rowptr"], cora["colidx"])
    assert np.array_equal(rp, g.rowptr) and np.array_equal(ci, g.colidx)


def test_partition1d_h_bit_exact(abi, golden, cora, small_graph):
    from graphaibench_b200 import ops
    for name, (rp, ci) in (("cora", (cora["rowptr64"], cora["colidx"])), ("sg", (small_graph["rowptr64"], small_graph["colidx"]))):
        for nparts in (2, 4):
            for part in range(nparts):
                r = ops.partition1d(rp, ci, nparts, part)
                key = f"part_{name}_{nparts}_{part}"
                assert list(golden[key + "_lb_le_m_ne"]) == [r["local_begin"], r["local_end"], len(r["idx_map"]), len(r["colidx"])]
                assert sha(r["idx_map"]) == str(golden[key + "_idx_sha"])
                assert sha(r["rowptr"]) == str(golden[key + "_rowptr_sha"])
                assert sha(r["colidx"]) == str(golden[key + "_colidx_sha"])
    # ragged: more parts than vertices, empty trailing partitions
    rp = np.array([0, 1, 2], np.int64); ci = np.array([1, 0], np.uint32)
    r = ops.partition1d(rp, ci, 4, 3)
    assert len(r["idx_map"]) == 0 and r["local_begin"] == r["local_end"] == 0


def test_argument_errors_are_reported_not_executed(abi):
    """Error behaviour of the C ABI (the reference asserts / exits; this layer returns GAI_ERR_ARG with a message and touches nothing).
    All of these are rejected before any CUDA call, so they run without a device."""
    L = abi.lib()
    ARG = 2
    h = C.c_void_p()
    rp = np.zeros(2, np.uint32)
    # NULL outputs / inputs
    assert L.gai_csr_create(1, 0, rp.ctypes.data_as(C.c_void_p), None, None, None) == ARG
    assert L.gai_csr_create(1, 0, None, None, None, C.byref(h)) == ARG
    # nnz >= 2^32: the 32-bit column-offset layout cannot address it (reference: eidType is 64-bit on disk, 32-bit in LearningGraph)
    assert L.gai_csr_create(1, 1 << 32, rp.ctypes.data_as(C.c_void_p), rp.ctypes.data_as(C.c_void_p), None, C.byref(h)) == ARG
    assert b"nnz" in L.gai_last_error()
    # rowptr[nv] must equal nnz
    assert L.gai_csr_create(1, 5, rp.ctypes.data_as(C.c_void_p), rp.ctypes.data_as(C.c_void_p), None, C.byref(h)) == ARG
    # NULL graph handles on every aggregation entry point
    assert L.gai_spmm_gcn(None, 4, None, 4, None, 4, 0, None, None) == ARG
    assert L.gai_spmm_mean(None, 4, None, 4, None, 4, 0, 0, None, None) == ARG
    assert L.gai_spmm_edge(None, 4, None, None, None, 4, None, 4, 0, None, None) == ARG
    assert L.gai_gat_forward_ld(None, 4, None, 4, None, None, 0.2, None, None, None, 4, 0, None) == ARG
    assert L.gai_csr_set_norms(None, None, None, None) == ARG
    assert L.gai_csr_set_row_segments(None, 0, None, None) == ARG
    assert L.gai_csr_build_transpose(None, None) == ARG
    # accessors on a NULL handle are total functions
    assert L.gai_csr_nv(None) == 0 and L.gai_csr_nnz(None) == 0 and L.gai_csr_rowptr(None) is None
    assert L.gai_csr_destroy(None) == 0 and L.gai_free(None) == 0
    # partition: bad part index / part count
    rp64 = np.array([0, 1, 2], np.int64); ci = np.array([1, 0], np.uint32)
    m, ne = C.c_int64(), C.c_int64()
    assert L.gai_partition1d_h(2, rp64.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), 0, 0, None, None, None, C.byref(m), C.byref(ne),
                               None, None) == ARG
    assert L.gai_partition1d_h(2, rp64.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), 2, 2, None, None, None, C.byref(m), C.byref(ne),
                               None, None) == ARG
    # out-parameters that must not be NULL
    assert L.gai_device_count(None) == ARG and L.gai_malloc(None, 16) == ARG and L.gai_event_create(None) == ARG


@pytest.mark.parametrize("name", ["r1_bench.json", "r2_bench.json"])
def test_committed_bench_line_has_the_contract_keys(name):
    """profiles/r{1,2}_bench.json are real `python bench.py` lines: the keys the driver and the judge read must all be there."""
    import json
    import os
    from conftest import ROOT
    d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["traffic"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert abs(d["value"] - d["config"]["csr_edges"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
