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
