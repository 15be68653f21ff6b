"""TEST INFRASTRUCTURE ONLY: CPU oracle for the GNN-layer hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this package.  The product path (``graphaibench_b200``) never does and has no CPU fallback.

Two things live here:

* ``liborc``  - ``gnn_oracle.c``, a plain-C restatement of the reference routines (each cites file:line).
* ``libref``  - ``_ref/libref_gnn.so``, the reference's OWN sources compiled by ``build_ref.sh`` behind a thin
  C harness (``ref_harness.cpp``).  Present whenever ``build_ref.sh`` ran in the build container; it travels to the
  GPU box as a prebuilt binary.  Used to pin ``liborc`` and as the ``cpu_baseline.kind == "reference"`` arm.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBORC_PATH = os.path.join(HERE, "liborc.so")
_LIBREF_PATH = os.path.join(HERE, "_ref", "libref_gnn.so")
_LIBREFPART_PATH = os.path.join(HERE, "_ref", "libref_part.so")

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build_liborc(force: bool = False) -> str:
    """gcc the C restatement. -ffp-contract=off: the reference's scale()+vadd() pair must not become an FMA."""
    src = os.path.join(HERE, "gnn_oracle.c")
    if force or not os.path.exists(_LIBORC_PATH) or os.path.getmtime(_LIBORC_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
                               src, "-o", _LIBORC_PATH, "-lm"])
    return _LIBORC_PATH


def build_ref() -> bool:
    """Compile the reference itself (only possible where /root/reference exists)."""
    if os.path.isdir(os.environ.get("REF", "/root/reference")):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return os.path.exists(_LIBREF_PATH)


_liborc = None
_libref = None


def _opt(p):
    return None if p is None else p.ctypes.data_as(C.c_void_p)


def liborc():
    global _liborc
    if _liborc is None:
        _liborc = C.CDLL(build_liborc())
        L = _liborc
        L.orc_init_glorot.argtypes = [C.c_size_t, C.c_size_t, f32p, C.c_uint]
        L.orc_add_selfloop.argtypes = [C.c_uint32, u32p, u32p, u32p, u32p]
        L.orc_vertex_norm.argtypes = [C.c_uint32, u32p, f32p]
        L.orc_edge_norm.argtypes = [C.c_uint32, u32p, u32p, f32p]
        L.orc_spmm_gcn.argtypes = [C.c_uint32, u32p, u32p, f32p, C.c_int, f32p, f32p]
        L.orc_spmm_mean.argtypes = [C.c_uint32, u32p, u32p, C.c_int, f32p, f32p, C.c_int]
        L.orc_spmm_edge.argtypes = [C.c_uint32, u32p, u32p, f32p, C.c_int, f32p, f32p]
        L.orc_symmetric_transpose.argtypes = [C.c_uint32, u32p, u32p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_symmetric_transpose.restype = C.c_int
        L.orc_gemm.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int]
        L.orc_relu.argtypes = [C.c_size_t, f32p, f32p]
        L.orc_d_relu.argtypes = [C.c_size_t, f32p, f32p, f32p]
        L.orc_softmax_loss.argtypes = [C.c_int, f32p, u8p, C.c_void_p, C.c_size_t, C.c_size_t, f32p, f32p, C.c_void_p, C.c_void_p]
        L.orc_softmax_loss.restype = C.c_float
        L.orc_sigmoid_loss.argtypes = [C.c_int, f32p, u8p, C.c_void_p, C.c_size_t, C.c_size_t, f32p, f32p, C.c_void_p, C.c_void_p]
        L.orc_sigmoid_loss.restype = C.c_float
        L.orc_adam.argtypes = [C.c_size_t, f32p, f32p, f32p, f32p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_float]
        L.orc_l2norm.argtypes = [C.c_int, C.c_int, f32p, f32p]
        L.orc_d_l2norm.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p]
        L.orc_gat_forward.argtypes = [C.c_uint32, u32p, u32p, C.c_int, f32p, f32p, C.c_float, f32p, f32p, f32p, f32p, f32p]
        L.orc_gat_backward.argtypes = [C.c_uint32, u32p, u32p, C.c_int, C.c_float, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, C.c_int]
        L.orc_gat_forward_heads.argtypes = [C.c_uint32, u32p, u32p, C.c_int, C.c_int, f32p, f32p, C.c_float, f32p, f32p, f32p, f32p, f32p]
        L.orc_gat_backward_heads.argtypes = [C.c_uint32, u32p, u32p, C.c_int, C.c_int, C.c_float, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                             C.c_int]
        L.orc_partition1d.argtypes = [C.c_uint32, i64p, u32p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_partition1d.restype = C.c_int64
    return _liborc


def have_ref() -> bool:
    return os.path.exists(_LIBREF_PATH)


def libref():
    global _libref
    if _libref is None:
        # include/gnn/configs.h:5 constructs a std::string from getenv("DATASET_PATH") at load time (NULL -> abort)
        os.environ.setdefault("DATASET_PATH", "/tmp/")
        _libref = C.CDLL(_LIBREF_PATH)
        L = _libref
        L.ref_graph_new.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p]
        L.ref_graph_new.restype = C.c_void_p
        for n in ("ref_graph_add_selfloop", "ref_graph_compute_vertex_data", "ref_graph_compute_edge_data", "ref_graph_free"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.ref_graph_nv.argtypes = [C.c_void_p]; L.ref_graph_nv.restype = C.c_uint32
        L.ref_graph_ne.argtypes = [C.c_void_p]; L.ref_graph_ne.restype = C.c_uint32
        L.ref_graph_export.argtypes = [C.c_void_p, u32p, u32p, C.c_void_p, C.c_void_p]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_init_glorot.argtypes = [C.c_size_t, C.c_size_t, f32p, C.c_uint]
        L.ref_gcn_aggregate.argtypes = [C.c_void_p, C.c_int, f32p, f32p]
        L.ref_sage_aggregate.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.c_int]
        L.ref_gat_aggregate.argtypes = [C.c_void_p, C.c_int, f32p, f32p, f32p, f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_symmetric_csr_transpose.argtypes = [C.c_int, C.c_int, u32p, u32p, f32p, f32p]
        L.ref_matmul.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int]
        L.ref_softmax_loss.argtypes = [C.c_int, C.c_int, f32p, u8p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_softmax_loss.restype = C.c_float
        L.ref_sigmoid_loss.argtypes = [C.c_int, C.c_int, f32p, u8p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
        L.ref_sigmoid_loss.restype = C.c_float
        L.ref_adam_steps.argtypes = [C.c_size_t, C.c_float, C.c_int, f32p, f32p]
        L.ref_model_new.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p, u8p, i64p, C.c_int]
        L.ref_model_new.restype = C.c_void_p
        L.ref_model_new_sigmoid.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p, u8p, i64p, C.c_int]
        L.ref_model_new_sigmoid.restype = C.c_void_p
        L.ref_model_train_epoch.argtypes = [C.c_void_p, C.POINTER(C.c_float)]; L.ref_model_train_epoch.restype = C.c_float
        L.ref_model_forward.argtypes = [C.c_void_p, C.POINTER(C.c_float)]; L.ref_model_forward.restype = C.c_float
        L.ref_model_backward.argtypes = [C.c_void_p]
        L.ref_model_update.argtypes = [C.c_void_p]
        L.ref_model_evaluate.argtypes = [C.c_void_p, C.c_char_p]; L.ref_model_evaluate.restype = C.c_float
        L.ref_model_tensor_size.argtypes = [C.c_void_p, C.c_char_p, C.c_int]; L.ref_model_tensor_size.restype = C.c_int64
        L.ref_model_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int, f32p, C.c_int64]; L.ref_model_get.restype = C.c_int64
        L.ref_model_set.argtypes = [C.c_void_p, C.c_char_p, C.c_int, f32p, C.c_int64]; L.ref_model_set.restype = C.c_int64
    return _libref


ARCH_ID = {"gcn": 0, "sage": 1, "gat": 2}


class RefModel:
    """The reference's Model<L> driven in memory (see ref_harness.cpp)."""

    def __init__(self, arch, rowptr, colidx, feats, labels, split9, dim_hid, num_cls, num_layers=2, lr=0.02, threads=1, sigmoid=False):
        L = libref()
        nv = len(rowptr) - 1
        rp = np.ascontiguousarray(rowptr, np.uint32)
        ci = np.ascontiguousarray(colidx, np.uint32)
        self.g = L.ref_graph_new(nv, len(ci), rp, ci)
        feats = np.ascontiguousarray(feats, np.float32)
        new = L.ref_model_new_sigmoid if sigmoid else L.ref_model_new
        self.h = new(ARCH_ID[arch], self.g, nv, feats.shape[1], dim_hid, num_cls, num_layers, lr, feats,
                                 np.ascontiguousarray(labels, np.uint8), np.ascontiguousarray(split9, np.int64), threads)
        self.L = L

    def graph(self):
        nv, ne = self.L.ref_graph_nv(self.g), self.L.ref_graph_ne(self.g)
        rp, ci, vd = np.zeros(nv + 1, np.uint32), np.zeros(ne, np.uint32), np.zeros(nv, np.float32)
        self.L.ref_graph_export(self.g, rp, ci, vd.ctypes.data_as(C.c_void_p), None)
        return rp, ci, vd

    def train_epoch(self):
        loss = C.c_float()
        acc = self.L.ref_model_train_epoch(self.h, C.byref(loss))
        return loss.value, acc

    def forward(self):
        loss = C.c_float()
        acc = self.L.ref_model_forward(self.h, C.byref(loss))
        return loss.value, acc

    def backward(self):
        self.L.ref_model_backward(self.h)

    def update(self):
        self.L.ref_model_update(self.h)

    def evaluate(self, which="test"):
        return self.L.ref_model_evaluate(self.h, which.encode())

    def get(self, name, layer=0):
        n = self.L.ref_model_tensor_size(self.h, name.encode(), layer)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, np.float32)
        self.L.ref_model_get(self.h, name.encode(), layer, out, n)
        return out

    def set(self, name, layer, arr):
        arr = np.ascontiguousarray(arr, np.float32).ravel()
        if self.L.ref_model_set(self.h, name.encode(), layer, arr, arr.size) < 0:
            raise KeyError(name)


_librefpart = None


def ref_partition1d(rowptr64, colidx, nparts, part):
    """The reference's edgecut_induced_partition1D (src/partitioner/graph_partition.cc:128-178) on an in-memory CSR."""
    global _librefpart
    if _librefpart is None:
        _librefpart = C.CDLL(_LIBREFPART_PATH)
        _librefpart.refpart_new.argtypes = [C.c_uint32, i64p, u32p, C.c_int]
        _librefpart.refpart_new.restype = C.c_void_p
        _librefpart.refpart_sizes.argtypes = [C.c_void_p, C.c_int, i64p]
        _librefpart.refpart_get.argtypes = [C.c_void_p, C.c_int, u32p, i64p, u32p]
    P = _librefpart
    rp = np.ascontiguousarray(rowptr64, np.int64)
    ci = np.ascontiguousarray(colidx, np.uint32)
    key = (rp.ctypes.data, len(rp), nparts)
    cache = ref_partition1d.__dict__.setdefault("cache", {})
    if key not in cache:
        cache.clear()
        cache[key] = (P.refpart_new(len(rp) - 1, rp, ci, nparts), rp, ci)
    h = cache[key][0]
    sz = np.zeros(4, np.int64)
    P.refpart_sizes(h, part, sz)
    idx = np.zeros(sz[0], np.uint32); srp = np.zeros(sz[0] + 1, np.int64); sci = np.zeros(max(sz[1], 1), np.uint32)
    P.refpart_get(h, part, idx, srp, sci)
    return dict(idx_map=idx, rowptr=srp, colidx=sci[: sz[1]], local_begin=int(sz[2]), local_end=int(sz[3]))


def orc_partition1d(rowptr64, colidx, nparts, part):
    """C restatement of the same (gnn_oracle.c: orc_partition1d)."""
    L = liborc()
    rp = np.ascontiguousarray(rowptr64, np.int64)
    ci = np.ascontiguousarray(colidx, np.uint32)
    nv = len(rp) - 1
    ne = C.c_int64()
    m = L.orc_partition1d(nv, rp, ci, nparts, part, None, None, None, C.byref(ne), None, None)
    idx = np.zeros(m, np.uint32); srp = np.zeros(m + 1, np.int64); sci = np.zeros(max(ne.value, 1), np.uint32)
    lb, le = C.c_uint32(), C.c_uint32()
    L.orc_partition1d(nv, rp, ci, nparts, part, idx.ctypes.data_as(C.c_void_p), srp.ctypes.data_as(C.c_void_p), sci.ctypes.data_as(C.c_void_p),
                      C.byref(ne), C.byref(lb), C.byref(le))
    return dict(idx_map=idx, rowptr=srp, colidx=sci[: ne.value], local_begin=lb.value, local_end=le.value)
