"""TEST INFRASTRUCTURE ONLY: the reference's Model<L>/layer orchestration restated over liborc (numpy buffers).

Mirrors, with citations into /root/reference:
  * Model::construct_network / forward_prop / backward_prop / update_weights / evaluate   src/gnn/net.cpp:422-615
  * GCN_layer / SAGE_layer / GAT_layer forward, backward, update_weight                   src/gnn/gconv/*_layer.cpp
  * dense_layer, l2norm_layer, softmax_loss_layer                                         src/layers/*.cpp
  * adam bookkeeping (shared vs per-layer optimiser objects)                              SURVEY.md §8 row A11
Dropout is not restated (every config runs rate 0 and the reference's masks are seeded from /dev/urandom).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import liborc


def _f(a):
    return np.ascontiguousarray(a, np.float32)


class Graph:
    """LearningGraph (include/gnn/lgraph.h:20-273) as numpy arrays: u32 rowptr/colidx + vertex_data_."""

    def __init__(self, rowptr, colidx):
        self.rowptr = np.ascontiguousarray(rowptr, np.uint32)
        self.colidx = np.ascontiguousarray(colidx, np.uint32)
        self.vdata = None
        self._perm = None

    @property
    def nv(self):
        return len(self.rowptr) - 1

    @property
    def ne(self):
        return len(self.colidx)

    def add_selfloop(self):
        rp = np.zeros(self.nv + 1, np.uint32)
        ci = np.zeros(self.ne + self.nv, np.uint32)
        liborc().orc_add_selfloop(self.nv, self.rowptr, self.colidx, rp, ci)
        self.rowptr, self.colidx = rp, ci

    def compute_vertex_data(self):
        self.vdata = np.zeros(self.nv, np.float32)
        liborc().orc_vertex_norm(self.nv, self.rowptr, self.vdata)

    def transpose_perm(self):
        if self._perm is None:
            perm = np.zeros(self.ne, np.uint32)
            bad = liborc().orc_symmetric_transpose(self.nv, self.rowptr, self.colidx, None, None, perm.ctypes.data_as(C.c_void_p))
            assert bad == 0, "pattern not symmetric"
            self._perm = perm
        return self._perm


class Adam:
    """adam (include/utils/optimizer.h:99-116, src/utilities/optimizer.cpp:22-35). State is keyed by the weight
    buffer; b1_t/b2_t advance once per update() call on this object."""

    def __init__(self, lr):
        self.lr, self.b1, self.b2, self.eps = np.float32(lr), np.float32(0.9), np.float32(0.999), np.float32(1e-8)
        self.b1_t, self.b2_t = C.c_float(0.9), C.c_float(0.999)
        self.state = {}

    def update(self, dW, W):
        key = W.ctypes.data
        if key not in self.state:
            self.state[key] = (np.zeros(W.size, np.float32), np.zeros(W.size, np.float32))
        m, v = self.state[key]
        liborc().orc_adam(W.size, _f(dW).ravel(), W.reshape(-1), m, v, self.lr, self.b1, self.b2, C.byref(self.b1_t), C.byref(self.b2_t), self.eps)


def glorot(dx, dy, seed):
    w = np.zeros(dx * dy, np.float32)
    liborc().orc_init_glorot(dx, dy, w, seed)
    return w


def matmul(x, y, z, A, B, Cm, ta=False, tb=False, accum=False):
    liborc().orc_gemm(x, y, z, A.reshape(-1), B.reshape(-1), Cm.reshape(-1), int(ta), int(tb), int(accum))


class ConvLayer:
    """graph_conv_layer<A> (include/layers/graph_conv_layer.h:6-61, src/gnn/graph_conv_layer.cpp:4-51)."""

    def __init__(self, arch, level, nv, din, dout, graph, act, lr, heads=1):
        self.arch, self.level, self.nv, self.din, self.dout, self.g, self.act = arch, level, nv, din, dout, graph, act
        self.heads = heads  # GAT extension (the reference has one head): see orc_gat_forward_heads
        self.W = glorot(din, dout, 1)
        self.W_grad = np.zeros(din * dout, np.float32)
        if arch == "sage":
            self.W_self = glorot(din, dout, 2)
            self.W_self_grad = np.zeros(din * dout, np.float32)
        self.in_temp = np.zeros(nv * din, np.float32)
        self.out_temp = np.zeros(nv * dout, np.float32)
        self.in_temp1 = np.zeros(nv * din, np.float32) if din <= dout else None
        self.feat_in = np.zeros(nv * din, np.float32) if level > 0 else None
        self.grad_in = np.zeros(nv * dout, np.float32)
        self.optm = Adam(lr)
        if arch == "gat":  # GAT_Aggregator::init, src/gnn/gconv/gat_aggregator.cpp:3-24
            ne = graph.ne
            self.alpha_l, self.alpha_r = glorot(dout, 1, 2), glorot(dout, 1, 3)
            self.alpha_lgrad, self.alpha_rgrad = np.zeros(dout, np.float32), np.zeros(dout, np.float32)
            self.scores, self.temp_scores = np.zeros(ne * heads, np.float32), np.zeros(ne * heads, np.float32)
            self.norm_scores, self.norm_scores_grad = np.zeros(ne * heads, np.float32), np.zeros(ne * heads, np.float32)
            self.alpha_opt = Adam(lr)
            self.closed_form = False

    # -- aggregators ---------------------------------------------------------------------------------
    def _aggregate(self, length, x, out, transposed=False):
        g, L = self.g, liborc()
        if self.arch == "gcn":  # symmetric: d_aggregate == aggregate (gcn_aggregator.cpp:23-46)
            L.orc_spmm_gcn(g.nv, g.rowptr, g.colidx, g.vdata, length, x, out)
        else:
            L.orc_spmm_mean(g.nv, g.rowptr, g.colidx, length, x, out, int(transposed))

    # -- forward -------------------------------------------------------------------------------------
    def forward(self, feat_out):
        x, y, z = self.nv, self.din, self.dout
        L = liborc()
        if self.arch == "gat":  # gat_layer.cpp:3-22
            matmul(x, z, y, self.feat_in, self.W, self.out_temp)
            g = self.g
            if self.heads == 1:
                L.orc_gat_forward(g.nv, g.rowptr, g.colidx, z, self.alpha_l, self.alpha_r, 0.2, self.out_temp,
                                  self.temp_scores, self.scores, self.norm_scores, feat_out)
            else:
                L.orc_gat_forward_heads(g.nv, g.rowptr, g.colidx, z, self.heads, self.alpha_l, self.alpha_r, 0.2, self.out_temp,
                                        self.temp_scores, self.scores, self.norm_scores, feat_out)
        else:  # gcn_layer.cpp:5-28, sage_layer.cpp:5-26
            if y > z:
                matmul(x, z, y, self.feat_in, self.W, self.out_temp)
                self._aggregate(z, self.out_temp, feat_out)
            else:
                self._aggregate(y, self.feat_in, self.in_temp1)
                matmul(x, z, y, self.in_temp1, self.W, feat_out)
            if self.arch == "sage":
                matmul(x, z, y, self.feat_in, self.W_self, feat_out, accum=True)
        if self.act:
            L.orc_relu(x * z, feat_out, feat_out)

    # -- backward ------------------------------------------------------------------------------------
    def backward(self, feat_out, grad_out):
        x, y, z = self.nv, self.din, self.dout
        L = liborc()
        if self.act:
            L.orc_d_relu(x * z, self.grad_in, feat_out, self.grad_in)
        if self.arch == "gat":  # gat_layer.cpp:24-42 (d_aggregate writes dZ over out_temp)
            g = self.g
            dz = np.zeros(x * z, np.float32)
            if self.heads == 1:
                L.orc_gat_backward(g.nv, g.rowptr, g.colidx, z, 0.2, self.out_temp, self.grad_in, self.temp_scores, self.norm_scores,
                                   self.scores, self.norm_scores_grad, self.alpha_lgrad, self.alpha_rgrad, dz, int(self.closed_form))
            else:
                L.orc_gat_backward_heads(g.nv, g.rowptr, g.colidx, z, self.heads, 0.2, self.out_temp, self.grad_in, self.temp_scores,
                                         self.norm_scores, self.scores, self.norm_scores_grad, self.alpha_lgrad, self.alpha_rgrad, dz,
                                         int(self.closed_form))
            self.out_temp[:] = dz
            if self.level != 0:
                matmul(x, y, z, self.out_temp, self.W, grad_out, False, True)
            matmul(y, z, x, self.feat_in, self.out_temp, self.W_grad, True, False)
            return
        if self.arch == "sage":  # sage_layer.cpp:37
            matmul(y, z, x, self.feat_in, self.grad_in, self.W_self_grad, True, False)
        if y > z:
            self._aggregate(z, self.grad_in, self.out_temp, transposed=True)
            if self.level > 0:
                matmul(x, y, z, self.out_temp, self.W, grad_out, False, True)
            matmul(y, z, x, self.feat_in, self.out_temp, self.W_grad, True, False)
        else:
            if self.level > 0:
                matmul(x, y, z, self.grad_in, self.W, self.in_temp, False, True)
                self._aggregate(y, self.in_temp, grad_out, transposed=True)
            matmul(y, z, x, self.in_temp1, self.grad_in, self.W_grad, True, False)
        if self.arch == "sage" and self.level > 0:  # sage_layer.cpp:50
            matmul(x, y, z, self.grad_in, self.W_self, grad_out, False, True, True)

    def update_weight(self, shared_opt):
        if self.arch == "gcn":  # gcn_layer.cpp:62-66: the Model's shared optimiser
            shared_opt.update(self.W_grad, self.W)
        elif self.arch == "sage":  # sage_layer.cpp:55-59: the layer's own optimiser, two calls
            self.optm.update(self.W_grad, self.W)
            self.optm.update(self.W_self_grad, self.W_self)
        else:  # gat_layer.cpp:44-48 + gat_aggregator.cpp:202-205
            shared_opt.update(self.W_grad, self.W)
            self.alpha_opt.update(self.alpha_lgrad, self.alpha_l)
            self.alpha_opt.update(self.alpha_rgrad, self.alpha_r)


class OracleModel:
    """Model<L> for subg_size == 0, softmax loss (src/gnn/net.cpp)."""

    def __init__(self, arch, rowptr, colidx, feats, labels, split9, dim_hid, num_cls, num_layers=2, lr=0.02, sigmoid=False, heads=1):
        self.arch = arch
        self.sigmoid = sigmoid  # argv[4] == "sigmoid": multi-hot labels, sigmoid_loss_layer, micro-F1 as accuracy (net.cpp:20,447-451)
        self.g = Graph(rowptr, colidx)
        if arch != "sage":
            self.g.add_selfloop()  # net.cpp:96
        self.g.compute_vertex_data()  # net.cpp:199-202
        nv = self.g.nv
        self.nv, self.ncls, self.hid, self.nl = nv, num_cls, dim_hid, num_layers
        self.feats = _f(feats).reshape(-1)
        dim_init = np.asarray(feats).shape[1]
        self.labels = np.ascontiguousarray(labels, np.uint8)
        if sigmoid:  # Reader::bin_read_vlabels(labels, false), reader.cpp:347-412
            hot = np.zeros((len(self.labels), num_cls), np.uint8)
            ok = self.labels < num_cls
            hot[np.nonzero(ok)[0], self.labels[ok]] = 1
            self.labels = np.ascontiguousarray(hot.reshape(-1))
        (self.tb, self.te, self.tc, self.vb, self.ve, self.vc, self.sb, self.se, self.sc) = [int(v) for v in split9]
        self.masks = {}
        for name, (b, e) in {"train": (self.tb, self.te), "val": (self.vb, self.ve), "test": (self.sb, self.se)}.items():
            m = np.zeros(nv, np.uint8); m[b:e] = 1; self.masks[name] = m
        self.use_dense = self.use_l2norm = arch == "gat"  # net.cpp:67-71
        self.layers = []
        for l in range(num_layers - 1):  # net.cpp:426-430
            self.layers.append(ConvLayer(arch, l, nv, dim_init if l == 0 else dim_hid, dim_hid, self.g, True, lr, heads))
        dim_out = dim_hid if self.use_dense else num_cls
        self.layers.append(ConvLayer(arch, num_layers - 1, nv, dim_hid, dim_out, self.g, False, lr, heads))
        self.layers[0].feat_in = self.feats
        if self.use_l2norm:
            self.l2_feat_in = np.zeros(nv * dim_hid, np.float32); self.l2_grad_in = np.zeros(nv * dim_hid, np.float32)
        if self.use_dense:  # dense_layer.cpp:4-40
            self.d_feat_in = np.zeros(nv * dim_hid, np.float32); self.d_grad_in = np.zeros(nv * num_cls, np.float32)
            self.d_W = glorot(dim_hid, num_cls, 1); self.d_W_grad = np.zeros(dim_hid * num_cls, np.float32)
            self.d_opt = Adam(lr)
        self.logits = np.zeros(nv * num_cls, np.float32)
        self.probs = np.zeros(nv * num_cls, np.float32)
        self.losses = np.zeros(nv, np.float32)
        self.opt = Adam(lr)  # net.cpp:362

    def _forward_layers(self):  # net.cpp:457-471
        Ls = self.layers
        for l in range(self.nl - 1):
            Ls[l].forward(Ls[l + 1].feat_in)
        if self.use_dense:
            Ls[-1].forward(self.l2_feat_in)
            liborc().orc_l2norm(self.nv, self.hid, self.l2_feat_in, self.d_feat_in)
            matmul(self.nv, self.ncls, self.hid, self.d_feat_in, self.d_W, self.logits)
        else:
            Ls[-1].forward(self.logits)

    def _loss(self, mask, b, e, grad):
        fn = liborc().orc_sigmoid_loss if self.sigmoid else liborc().orc_softmax_loss
        acc = C.c_float()
        loss = fn(self.ncls, self.logits, self.labels, mask.ctypes.data_as(C.c_void_p), b, e, self.probs, self.losses,
                  grad.ctypes.data_as(C.c_void_p) if grad is not None else None, C.byref(acc))
        return float(loss), float(acc.value)

    def forward(self):
        self._forward_layers()
        return self._loss(self.masks["train"], self.tb, self.te, None)

    def backward(self):  # net.cpp:580-615
        Ls = self.layers
        last_grad = self.d_grad_in if self.use_dense else Ls[-1].grad_in
        self._loss(self.masks["train"], self.tb, self.te, last_grad)
        if self.use_dense:  # dense_layer.cpp:57-72 (updates its own weights inside backward), l2norm_layer.cpp:40-64
            matmul(self.hid, self.ncls, self.nv, self.d_feat_in, self.d_grad_in, self.d_W_grad, True)
            matmul(self.nv, self.hid, self.ncls, self.d_grad_in, self.d_W, self.l2_grad_in, False, True)
            self.d_opt.update(self.d_W_grad, self.d_W)
            liborc().orc_d_l2norm(self.nv, self.hid, self.l2_feat_in, self.l2_grad_in, Ls[-1].grad_in)
            Ls[-1].backward(self.l2_feat_in, Ls[-2].grad_in)
        else:
            Ls[-1].backward(self.logits, Ls[-2].grad_in)
        for l in range(self.nl - 2, 0, -1):
            Ls[l].backward(Ls[l + 1].feat_in, Ls[l - 1].grad_in)
        Ls[0].backward(Ls[1].feat_in, None)

    def update(self):  # net.cpp:230-234
        for y in self.layers:
            y.update_weight(self.opt)

    def train_epoch(self):
        out = self.forward()
        self.backward()
        self.update()
        return out

    def evaluate(self, which="test"):  # net.cpp:506-577 (softmax branch: argmax over logits only)
        self._forward_layers()
        b, e = (self.vb, self.ve) if which == "val" else (self.sb, self.se)
        if self.sigmoid:  # net.cpp:569-572: loss forward over the range, then micro-F1 of the sigmoid outputs
            return self._loss(self.masks["val" if which == "val" else "test"], b, e, None)[1]
        lg = self.logits.reshape(self.nv, self.ncls)[b:e]
        pred = np.argmax(lg, axis=1)
        return float(np.float32(np.sum(pred == self.labels[b:e])) / np.float32(e - b))


def read_dataset(path):
    """Reader::bin_read_* (src/gnn/reader.cpp:248-457): graph.meta.txt + graph.vertex.bin (int64) + graph.edge.bin (u32)
    + graph.feats.bin (f32 row-major) + graph.vlabel.bin (u8).  Returns dict of numpy arrays."""
    import os
    meta = [int(t) for t in open(os.path.join(path, "graph.meta.txt")).read().split()]
    nv, ne, feat_len, ncls = meta[0], meta[1], meta[7], meta[8]
    split = meta[10:19] if len(meta) >= 19 else None
    rp64 = np.fromfile(os.path.join(path, "graph.vertex.bin"), np.int64, nv + 1)
    ci = np.fromfile(os.path.join(path, "graph.edge.bin"), np.uint32, ne)
    out = dict(nv=nv, ne=ne, feat_len=feat_len, ncls=ncls, split=np.array(split, np.int64) if split else None,
               rowptr64=rp64, rowptr=rp64.astype(np.uint32), colidx=ci, max_degree=meta[6])
    fp = os.path.join(path, "graph.feats.bin")
    if feat_len and os.path.exists(fp):
        out["feats"] = np.fromfile(fp, np.float32, nv * feat_len).reshape(nv, feat_len)
    lp = os.path.join(path, "graph.vlabel.bin")
    if os.path.exists(lp):
        out["labels"] = np.fromfile(lp, np.uint8, nv)
    return out
