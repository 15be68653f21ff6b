"""ctypes front end over libgai_host.so: the C++ Model<L> / layer classes (graphaibench_b200/host) that mirror the
reference's include/gnn/net.h API. Python only carries pointers; every numeric step runs in the sm_100a kernels."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi

HOSTLIB_PATH = os.path.join(_abi.PKG, "libgai_host.so")
ARCH_ID = {"gcn": 0, "sage": 1, "gat": 2}
_h = None


def hostlib():
    global _h
    if _h is None:
        _abi.lib()  # resolves libgai_b200.so first
        if not os.path.exists(HOSTLIB_PATH):
            raise OSError(f"{HOSTLIB_PATH} not built: run `python -m graphaibench_b200.build`")
        L = C.CDLL(HOSTLIB_PATH)
        L.gai_host_set_stream.argtypes = [C.c_void_p]
        L.gai_graph_new.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.gai_graph_new.restype = C.c_void_p
        L.gai_model_new.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gai_model_new.restype = C.c_void_p
        for n in ("gai_model_train_epoch", "gai_model_forward"):
            getattr(L, n).argtypes = [C.c_void_p, C.POINTER(C.c_float)]
            getattr(L, n).restype = C.c_float
        L.gai_model_backward.argtypes = [C.c_void_p]
        L.gai_model_update.argtypes = [C.c_void_p]
        L.gai_model_evaluate.argtypes = [C.c_void_p, C.c_char_p]
        L.gai_model_evaluate.restype = C.c_float
        L.gai_model_refresh_inputs.argtypes = [C.c_void_p, C.c_void_p]
        L.gai_model_prefetch_features.argtypes = [C.c_void_p, C.c_void_p]
        L.gai_model_tensor_size.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.gai_model_tensor_size.restype = C.c_int64
        L.gai_model_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]
        L.gai_model_get.restype = C.c_int64
        L.gai_model_set.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]
        L.gai_model_set.restype = C.c_int64
        L.gai_host_profile_enable.argtypes = [C.c_int]
        L.gai_host_profile_json.argtypes = [C.c_char_p, C.c_int64]
        L.gai_host_profile_json.restype = C.c_int64
        L.gai_model_sync.argtypes = []
        L.gai_host_glorot.argtypes = [C.c_uint64, C.c_uint64, C.c_uint, C.c_void_p]
        L.gai_comm_new.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.gai_comm_new.restype = C.c_void_p
        L.gai_comm_free.argtypes = [C.c_void_p]
        L.gai_comm_barrier.argtypes = [C.c_void_p]
        L.gai_comm_check.argtypes = [C.c_void_p]
        L.gai_model_new_partitioned.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                C.c_void_p, C.c_void_p, C.c_void_p]
        L.gai_model_new_partitioned.restype = C.c_void_p
        L.gai_model_halo_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.gai_host_train_partitioned.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gai_host_train_partitioned.restype = C.c_int
        L.gai_host_partition_rows.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gai_host_partition_rows.restype = C.c_int
        _h = L
    return _h


def glorot(dim_x, dim_y, seed):
    """Reference-exact Glorot-uniform initial weights (init_glorot, math_functions.cpp:11-19), [dim_x, dim_y] fp32 on the host."""
    out = np.empty((dim_x, dim_y), np.float32)
    hostlib().gai_host_glorot(dim_x, dim_y, seed, out.ctypes.data_as(C.c_void_p))
    return out


def profile_enable(on: bool):
    hostlib().gai_host_profile_enable(int(on))


def profile_collect():
    """Per-(bucket, shape) device timings recorded since profile_enable(True): list of dicts (calls, ms, bytes, flops)."""
    import json
    buf = C.create_string_buffer(1 << 20)
    hostlib().gai_host_profile_json(buf, len(buf))
    return json.loads(buf.value.decode())


class GnnModel:
    """Model<GCN_layer|SAGE_layer|GAT_layer> built from in-memory arrays (raw graph: no self-loops; the model adds them
    for GCN/GAT exactly as Model::load_data does)."""

    def __init__(self, arch, rowptr, colidx, feats, labels, split9, dim_hid, num_cls, num_layers=2, lr=0.02, stream=None):
        L = hostlib()
        L.gai_host_set_stream(C.c_void_p(stream) if stream else None)
        rp = np.ascontiguousarray(rowptr, np.uint32)
        ci = np.ascontiguousarray(colidx, np.uint32)
        self.feats = np.ascontiguousarray(feats, np.float32)
        labels = np.ascontiguousarray(labels, np.uint8)
        split = np.ascontiguousarray(split9, np.int64)
        self.nv, self.dim_init = self.feats.shape
        g = L.gai_graph_new(self.nv, len(ci), rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p))
        self.h = L.gai_model_new(ARCH_ID[arch], g, self.dim_init, dim_hid, num_cls, num_layers, lr, self.feats.ctypes.data_as(C.c_void_p),
                                 labels.ctypes.data_as(C.c_void_p), split.ctypes.data_as(C.c_void_p))
        if not self.h:
            raise ValueError(arch)
        self.L = L

    def train_epoch(self):
        loss = C.c_float()
        acc = self.L.gai_model_train_epoch(self.h, C.byref(loss))
        return loss.value, acc

    def forward(self):
        loss = C.c_float()
        acc = self.L.gai_model_forward(self.h, C.byref(loss))
        return loss.value, acc

    def backward(self):
        self.L.gai_model_backward(self.h)

    def update(self):
        self.L.gai_model_update(self.h)

    def evaluate(self, which="test"):
        return self.L.gai_model_evaluate(self.h, which.encode())

    def refresh_inputs(self, feats_host_ptr=None):
        self.L.gai_model_refresh_inputs(self.h, feats_host_ptr)

    def prefetch_inputs(self, feats_host_ptr=None):
        """Start the next step's host->device feature copy on a copy stream (swapped in by the next refresh_inputs())."""
        self.L.gai_model_prefetch_features(self.h, feats_host_ptr)

    def get(self, name, layer=0):
        n = self.L.gai_model_tensor_size(self.h, name.encode(), layer)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, np.float32)
        self.L.gai_model_get(self.h, name.encode(), layer, out.ctypes.data_as(C.c_void_p), n)
        return out

    def set(self, name, layer, arr):
        arr = np.ascontiguousarray(arr, np.float32).ravel()
        if self.L.gai_model_set(self.h, name.encode(), layer, arr.ctypes.data_as(C.c_void_p), arr.size) < 0:
            raise KeyError(name)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def owner_range(nv, world, rank):
    """graph_partition.cc:131-140: S = ceil(nv / P); rank p owns [p*S, min((p+1)*S, nv))."""
    S = (nv + world - 1) // world
    first = min(S * rank, nv)
    return S, first, min(first + S, nv)


def partition_rows(world, rank, nv_global, rows_rowptr, rows_colidx, selfloops=False):
    """Host-only integer part of the 1D partition (LearningGraph::partition_rows): (rowptr u32, local colidx u32, halo global ids u32)."""
    L = hostlib()
    rp = np.ascontiguousarray(rows_rowptr, np.int64)
    ci = np.ascontiguousarray(rows_colidx, np.uint32)
    sizes = np.zeros(2, np.uint64)
    L.gai_host_partition_rows(world, rank, nv_global, _vp(rp), _vp(ci), int(selfloops), _vp(sizes), None, None, None)
    rp_out = np.zeros(len(rp), np.uint32); ci_out = np.zeros(max(int(sizes[1]), 1), np.uint32); halo = np.zeros(max(int(sizes[0]), 1), np.uint32)
    L.gai_host_partition_rows(world, rank, nv_global, _vp(rp), _vp(ci), int(selfloops), _vp(sizes), _vp(rp_out), _vp(ci_out), _vp(halo))
    return rp_out, ci_out[: int(sizes[1])], halo[: int(sizes[0])]


def train_partitioned_inprocess(arch, world, rowptr64, colidx, feats, labels, split9, dim_hid, num_cls, num_layers=2, lr=0.02, epochs=3):
    """`world` ranks as host threads of this process (rank r on device r mod device count): the whole C++ partitioned path — partition,
    peer registration, halo pulls, weight-gradient and statistics combination — on whatever devices are visible, one included.
    Returns dict(losses, accs, test_acc, weights (flat: layers in order, W then W_self), halo (world x {masters, halo, exchanges, bytes}))."""
    L = hostlib()
    rp = np.ascontiguousarray(rowptr64, np.int64); ci = np.ascontiguousarray(colidx, np.uint32)
    fe = np.ascontiguousarray(feats, np.float32); la = np.ascontiguousarray(labels, np.uint8); sp = np.ascontiguousarray(split9, np.int64)
    nv, F = fe.shape
    dims = [F] + [dim_hid] * (num_layers - 1) + [num_cls]
    nw = sum(dims[l] * dims[l + 1] for l in range(num_layers)) * (2 if arch == "sage" else 1)
    losses, accs, tacc = np.zeros(epochs, np.float32), np.zeros(epochs, np.float32), np.zeros(1, np.float32)
    w = np.zeros(nw, np.float32); halo = np.zeros((world, 4), np.uint64)
    rc = L.gai_host_train_partitioned(ARCH_ID[arch], world, nv, _vp(rp), _vp(ci), F, dim_hid, num_cls, num_layers, lr, _vp(fe), _vp(la), _vp(sp), epochs,
                                      _vp(losses), _vp(accs), _vp(tacc), _vp(w), _vp(halo))
    if rc != 0:
        raise RuntimeError("gai_host_train_partitioned failed (no CUDA device?)")
    return dict(losses=losses, accs=accs, test_acc=float(tacc[0]), weights=w, halo=halo)


ALLGATHER_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def torch_allgather_callback(group=None, device=None):
    """The bootstrap all-gather of gai_peers_create over torch.distributed (NCCL: staged through `device`; gloo: CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)

    def fn(_ctx, send, nbytes, recv_all):
        buf = (C.c_ubyte * nbytes).from_address(send)
        t = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        if device is not None:
            t = t.to(device)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        flat = torch.cat(outs).cpu().numpy().tobytes()
        C.memmove(recv_all, flat, len(flat))
    return ALLGATHER_FN(fn)


class DistGnnModel(GnnModel):
    """One rank of the C++ partitioned Model<GCN_layer | SAGE_layer> (host/gai_model.h: set_comm + init_partitioned). The caller passes this
    rank's rows of the RAW graph (rowptr rebased to 0, GLOBAL column ids), its rows of features / labels and the GLOBAL split ranges."""

    def __init__(self, arch, rank, world, allgather_cb, nv_global, rows_rowptr, rows_colidx, feats_local, labels_local, split9_global, dim_hid,
                 num_cls, num_layers=2, lr=0.02, stream=None):
        L = hostlib()
        L.gai_host_set_stream(C.c_void_p(stream) if stream else None)
        self._cb = allgather_cb  # keep the ctypes callback alive
        self.comm = L.gai_comm_new(rank, world, C.cast(allgather_cb, C.c_void_p), None)
        rp = np.ascontiguousarray(rows_rowptr, np.int64); ci = np.ascontiguousarray(rows_colidx, np.uint32)
        self.feats = np.ascontiguousarray(feats_local, np.float32)
        la = np.ascontiguousarray(labels_local, np.uint8); sp = np.ascontiguousarray(split9_global, np.int64)
        self.nv, self.dim_init = self.feats.shape
        self.h = L.gai_model_new_partitioned(ARCH_ID[arch], self.comm, nv_global, _vp(rp), _vp(ci), self.dim_init, dim_hid, num_cls, num_layers, lr,
                                             _vp(self.feats), _vp(la), _vp(sp))
        if not self.h:
            raise ValueError(arch)
        self.L = L

    def halo_stats(self):
        out = np.zeros(4, np.uint64)
        self.L.gai_model_halo_stats(self.h, _vp(out))
        return dict(masters=int(out[0]), halo=int(out[1]), exchanges=int(out[2]), bytes=int(out[3]))

    def check(self):
        self.L.gai_comm_check(self.comm)
