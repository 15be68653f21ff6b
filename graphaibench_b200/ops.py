"""Torch-tensor front end over the C ABI (torch is plumbing: device memory + streams). Every function launches the
hand-written sm_100a kernels in libgai_b200.so on torch's current stream; nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._abi import check, lib

EPI_NONE, EPI_RELU, EPI_ADD, EPI_MASK, EPI_PADDED, EPI_BITMASK = 0, 1, 2, 4, 8, 16


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and (t.dim() == 0 or t.stride(-1) == 1), "device tensors with unit inner stride only"
    return C.c_void_p(t.data_ptr())


def _f32(t):
    assert t.dtype == torch.float32
    return _p(t)


class DeviceGraph:
    """gai_csr handle (device CSR + degree normalisers + hub list + cached transpose permutation)."""

    def __init__(self, rowptr, colidx, device_arrays=False):
        L = lib()
        self._h = C.c_void_p()
        if device_arrays:
            assert rowptr.dtype == torch.int32 or rowptr.dtype == torch.uint32
            self._keep = (rowptr, colidx)
            nv, nnz = rowptr.numel() - 1, colidx.numel()
            check(L.gai_csr_create_device(nv, nnz, _p(rowptr), _p(colidx), _stream(), C.byref(self._h)), "gai_csr_create_device")
        else:
            rp = np.ascontiguousarray(rowptr, np.uint32)
            ci = np.ascontiguousarray(colidx, np.uint32)
            nv, nnz = len(rp) - 1, len(ci)
            check(L.gai_csr_create(nv, nnz, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), _stream(), C.byref(self._h)),
                  "gai_csr_create")
        self.nv, self.nnz = nv, nnz

    def __del__(self):
        try:
            if self._h:
                lib().gai_csr_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def n_hub(self):
        return lib().gai_csr_num_hub_rows(self._h)

    def vertex_norm(self):
        out = torch.empty(self.nv, dtype=torch.float32, device="cuda")
        check(lib().gai_memcpy_d2d(_p(out), lib().gai_csr_vertex_norm(self._h), 4 * self.nv, _stream()))
        return out

    def csr(self):
        rp = torch.empty(self.nv + 1, dtype=torch.int32, device="cuda")
        ci = torch.empty(max(self.nnz, 1), dtype=torch.int32, device="cuda")
        check(lib().gai_memcpy_d2d(_p(rp), lib().gai_csr_rowptr(self._h), 4 * (self.nv + 1), _stream()))
        check(lib().gai_memcpy_d2d(_p(ci), lib().gai_csr_colidx(self._h), 4 * self.nnz, _stream()))
        return rp, ci[: self.nnz]

    def transpose_perm(self):
        check(lib().gai_csr_build_transpose(self._h, _stream()), "gai_csr_build_transpose")
        out = torch.empty(max(self.nnz, 1), dtype=torch.int32, device="cuda")
        check(lib().gai_memcpy_d2d(_p(out), lib().gai_csr_transpose_perm(self._h), 4 * self.nnz, _stream()))
        return out[: self.nnz]

    def set_row_segments(self, segments):
        """segments: list of (row_begin, row_end); the *_rows calls with exactly these bounds use a degree-ordered work list."""
        b = np.ascontiguousarray(np.asarray(segments, np.uint32).reshape(-1, 2))
        check(lib().gai_csr_set_row_segments(self._h, len(b), b.ctypes.data_as(C.c_void_p), _stream()), "gai_csr_set_row_segments")

    def set_norms(self, norm_gcn=None, norm_mean=None):
        check(lib().gai_csr_set_norms(self._h, _f32(norm_gcn) if norm_gcn is not None else None,
                                      _f32(norm_mean) if norm_mean is not None else None, _stream()))


def add_selfloop(rowptr, colidx):
    rp = np.ascontiguousarray(rowptr, np.uint32)
    ci = np.ascontiguousarray(colidx, np.uint32)
    nv = len(rp) - 1
    rpo = np.empty(nv + 1, np.uint32)
    cio = np.empty(len(ci) + nv, np.uint32)
    check(lib().gai_add_selfloop_h(nv, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), rpo.ctypes.data_as(C.c_void_p),
                                   cio.ctypes.data_as(C.c_void_p)), "gai_add_selfloop_h")
    return rpo, cio


def coo_to_csr(nv, src, dst, symmetrize=True):
    """Device CSR construction (csrc/convert.cu): int32/uint32 COO tensors on the GPU -> (rowptr int64[nv+1], colidx int32[nnz]) tensors:
    self-loops and ids >= nv dropped, mirrors added if `symmetrize`, duplicates removed, rows sorted — the reference converter's edge SET."""
    assert src.is_cuda and dst.is_cuda and src.numel() == dst.numel() and src.element_size() == 4 and dst.element_size() == 4
    src, dst = src.contiguous(), dst.contiguous()
    rp, ci, nnz = C.c_void_p(), C.c_void_p(), C.c_uint64()
    check(lib().gai_coo_to_csr(nv, src.numel(), src.data_ptr(), dst.data_ptr(), int(symmetrize), _stream(), C.byref(rp), C.byref(ci), C.byref(nnz)),
          "gai_coo_to_csr")
    rowptr = torch.empty(nv + 1, dtype=torch.int64, device=src.device)
    colidx = torch.empty(nnz.value, dtype=torch.int32, device=src.device)
    check(lib().gai_memcpy_d2d(rowptr.data_ptr(), rp, 8 * (nv + 1), _stream()), "gai_memcpy_d2d")
    if nnz.value:
        check(lib().gai_memcpy_d2d(colidx.data_ptr(), ci, 4 * nnz.value, _stream()), "gai_memcpy_d2d")
    check(lib().gai_stream_sync(_stream()), "gai_stream_sync")
    lib().gai_free(rp); lib().gai_free(ci)
    return rowptr, colidx


def add_selfloop_device(rowptr, colidx, first_id=0):
    """LearningGraph::add_selfloop on the device (int32 tensors): row r gains the id first_id + r at its sorted place."""
    nv = rowptr.numel() - 1
    rpo = torch.empty(nv + 1, dtype=torch.int32, device=rowptr.device)
    cio = torch.empty(colidx.numel() + nv, dtype=torch.int32, device=rowptr.device)
    check(lib().gai_add_selfloop_d(nv, first_id, rowptr.contiguous().data_ptr(), colidx.contiguous().data_ptr(), rpo.data_ptr(), cio.data_ptr(), _stream()),
          "gai_add_selfloop_d")
    return rpo, cio


def _ld(t):
    return t.stride(0) if t.dim() == 2 else t.numel()


def spmm_gcn(g, x, out=None, flags=0, addend=None, rows=None):
    F = x.shape[1]
    if out is None:
        out = torch.empty(g.nv, F, dtype=torch.float32, device=x.device)
    if rows is None:
        check(lib().gai_spmm_gcn(g.handle, F, _f32(x), x.stride(0), _f32(out), out.stride(0), flags, _p(addend), _stream()), "gai_spmm_gcn")
    else:
        check(lib().gai_spmm_gcn_rows(g.handle, rows[0], rows[1], F, _f32(x), x.stride(0), _f32(out), out.stride(0), flags, _p(addend), _stream()),
              "gai_spmm_gcn_rows")
    return out


def spmm_mean(g, x, out=None, transposed=False, flags=0, addend=None, rows=None):
    F = x.shape[1]
    if out is None:
        out = torch.empty(g.nv, F, dtype=torch.float32, device=x.device)
    if rows is None:
        check(lib().gai_spmm_mean(g.handle, F, _f32(x), x.stride(0), _f32(out), out.stride(0), int(transposed), flags, _p(addend), _stream()),
              "gai_spmm_mean")
    else:
        check(lib().gai_spmm_mean_rows(g.handle, rows[0], rows[1], F, _f32(x), x.stride(0), _f32(out), out.stride(0), int(transposed), flags,
                                       _p(addend), _stream()), "gai_spmm_mean_rows")
    return out


def spmm_edge(g, vals, x, out=None, perm=None, flags=0, addend=None):
    F = x.shape[1]
    if out is None:
        out = torch.empty(g.nv, F, dtype=torch.float32, device=x.device)
    check(lib().gai_spmm_edge(g.handle, F, _f32(vals), _p(perm), _f32(x), x.stride(0), _f32(out), out.stride(0), flags, _p(addend), _stream()),
          "gai_spmm_edge")
    return out


def matmul(A, B, out=None, transA=False, transB=False, accum=False, flags=0):
    """Reference matmul(x,y,z,A,B,C,transA,transB,accum): C[x,y] = op(A)[x,z] @ op(B)[z,y] (+C)."""
    x, z = (A.shape[1], A.shape[0]) if transA else (A.shape[0], A.shape[1])
    y = B.shape[0] if transB else B.shape[1]
    assert (B.shape[1] if transB else B.shape[0]) == z
    if out is None:
        assert not accum
        out = torch.empty(x, y, dtype=torch.float32, device=A.device)
    check(lib().gai_matmul_ld(x, y, z, _f32(A), A.stride(0), _f32(B), B.stride(0), _f32(out), out.stride(0), int(transA), int(transB),
                              int(accum), flags, _stream()), "gai_matmul_ld")
    return out


def matmul_kcat(A1, B1, A2, B2, out=None, transB=False, flags=0, mask=None, relu_bits=None):
    """C = A1·op(B1) + A2·op(B2) in one pass (gai_matmul_kcat); flags: EPI_RELU (relu_bits: int32 [x, ceil(y/32)] receives the sign bits of C),
    or EPI_MASK with `mask` (d_relu by the activation; with EPI_BITMASK `mask` is such a sign-bit tensor)."""
    x, z1, z2 = A1.shape[0], A1.shape[1], A2.shape[1]
    y = B1.shape[0] if transB else B1.shape[1]
    if out is None:
        out = torch.empty(x, y, dtype=torch.float32, device=A1.device)
    check(lib().gai_matmul_kcat(x, y, z1, _f32(A1), A1.stride(0), _f32(B1), B1.stride(0), z2, _f32(A2), A2.stride(0), _f32(B2), B2.stride(0),
                                _f32(out), out.stride(0), int(transB), flags, _p(mask), mask.stride(0) if mask is not None else 0,
                                _p(relu_bits), relu_bits.stride(0) if relu_bits is not None else 0, _stream()), "gai_matmul_kcat")
    return out


def matmul_relu_bits(A, B, relu_bits, out=None, flags=0):
    """C = ReLU(A·B) and the sign bits of C (gai_matmul_relu_bits)."""
    x, z = A.shape
    y = B.shape[1]
    if out is None:
        out = torch.empty(x, y, dtype=torch.float32, device=A.device)
    check(lib().gai_matmul_relu_bits(x, y, z, _f32(A), A.stride(0), _f32(B), B.stride(0), _f32(out), out.stride(0), flags, _p(relu_bits),
                                     relu_bits.stride(0), _stream()), "gai_matmul_relu_bits")
    return out


def matmul_mask(A, B, mask, out=None, transB=False, flags=0):
    """C = mask > 0 ? A·op(B) : 0 (gai_matmul_mask: the input gradient with the d_relu of the layer below folded in)."""
    x, z = A.shape
    y = B.shape[0] if transB else B.shape[1]
    if out is None:
        out = torch.empty(x, y, dtype=torch.float32, device=A.device)
    check(lib().gai_matmul_mask(x, y, z, _f32(A), A.stride(0), _f32(B), B.stride(0), _f32(out), out.stride(0), int(transB), _p(mask),
                                mask.stride(0), flags, _stream()), "gai_matmul_mask")
    return out


def matmul_ncat(A, B1, B2, out1=None, out2=None, flags=0):
    """C1 = A·B1, C2 = A·B2 with A read once (gai_matmul_ncat); flags: EPI_PADDED."""
    x, z = A.shape
    out1 = torch.empty(x, B1.shape[1], dtype=torch.float32, device=A.device) if out1 is None else out1
    out2 = torch.empty(x, B2.shape[1], dtype=torch.float32, device=A.device) if out2 is None else out2
    check(lib().gai_matmul_ncat(x, z, _f32(A), A.stride(0), B1.shape[1], _f32(B1), B1.stride(0), _f32(out1), out1.stride(0), B2.shape[1],
                                _f32(B2), B2.stride(0), _f32(out2), out2.stride(0), flags, _stream()), "gai_matmul_ncat")
    return out1, out2


def wgrad_two_a(A1, A2, B, out1=None, out2=None):
    """C1 = A1^T·B, C2 = A2^T·B with B read once (gai_wgrad_two_a)."""
    z, y = B.shape
    c1 = torch.empty(A1.shape[1], y, dtype=torch.float32, device=B.device) if out1 is None else out1
    c2 = torch.empty(A2.shape[1], y, dtype=torch.float32, device=B.device) if out2 is None else out2
    check(lib().gai_wgrad_two_a(z, y, _f32(B), B.stride(0), A1.shape[1], _f32(A1), A1.stride(0), _f32(c1), y, A2.shape[1], _f32(A2),
                                A2.stride(0), _f32(c2), y, _stream()), "gai_wgrad_two_a")
    return c1, c2


def wgrad_two_b(A, B1, B2, out1=None, out2=None):
    """C1 = A^T·B1, C2 = A^T·B2 with A read once (gai_wgrad_two_b)."""
    z, x = A.shape
    c1 = torch.empty(x, B1.shape[1], dtype=torch.float32, device=A.device) if out1 is None else out1
    c2 = torch.empty(x, B2.shape[1], dtype=torch.float32, device=A.device) if out2 is None else out2
    check(lib().gai_wgrad_two_b(z, x, _f32(A), A.stride(0), B1.shape[1], _f32(B1), B1.stride(0), _f32(c1), B1.shape[1], B2.shape[1],
                                _f32(B2), B2.stride(0), _f32(c2), B2.shape[1], _stream()), "gai_wgrad_two_b")
    return c1, c2


def dropout(x, rate, seed, call, out=None, mask=None):
    """out = x * mask / (1 - rate), mask ~ Bernoulli(1 - rate) from hash(seed, call, index) (gai_dropout). Returns (out, mask uint8)."""
    out = torch.empty_like(x) if out is None else out
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if mask is None else mask
    check(lib().gai_dropout(x.numel(), rate, 1.0 / (1.0 - rate), seed, call, _f32(x), _p(mask), _f32(out), _stream()), "gai_dropout")
    return out, mask


def d_dropout(grad, mask, rate, out=None):
    out = torch.empty_like(grad) if out is None else out
    check(lib().gai_d_dropout(grad.numel(), 1.0 / (1.0 - rate), _f32(grad), _p(mask), _f32(out), _stream()), "gai_d_dropout")
    return out


def relu(x, out=None):
    out = torch.empty_like(x) if out is None else out
    check(lib().gai_relu(x.numel(), _f32(x), _f32(out), _stream()), "gai_relu")
    return out


def d_relu(grad, data, out=None):
    out = torch.empty_like(grad) if out is None else out
    check(lib().gai_d_relu(grad.numel(), _f32(grad), _f32(data), _f32(out), _stream()), "gai_d_relu")
    return out


def l2norm(x, out=None):
    out = torch.empty_like(x) if out is None else out
    check(lib().gai_l2norm(x.shape[0], x.shape[1], _f32(x), _f32(out), _stream()), "gai_l2norm")
    return out


def d_l2norm(feat_in, grad_in, out=None):
    out = torch.empty_like(feat_in) if out is None else out
    check(lib().gai_d_l2norm(feat_in.shape[0], feat_in.shape[1], _f32(feat_in), _f32(grad_in), _f32(out), _stream()), "gai_d_l2norm")
    return out


def softmax_ce_forward(logits, labels, masks, begin, end, probs, losses):
    """logits / probs may be column views of row-padded buffers (their stride(0) is the row pitch)."""
    check(lib().gai_softmax_ce_forward_ld(logits.shape[1], begin, end, _p(masks), _p(labels), _f32(logits), logits.stride(0), _f32(probs),
                                          probs.stride(0), _f32(losses), _stream()), "gai_softmax_ce_forward_ld")


def softmax_ce_forward_stats(logits, labels, masks, begin, end, probs, losses, stats):
    """softmax_ce_forward + masked_loss_accuracy in one pass over the logits: stats = {mean loss, accuracy, count}."""
    check(lib().gai_softmax_ce_forward_stats_ld(logits.shape[1], begin, end, _p(masks), _p(labels), _f32(logits), logits.stride(0), _f32(probs),
                                                probs.stride(0), _f32(losses), _f32(stats), _stream()), "gai_softmax_ce_forward_stats_ld")
    return stats


def softmax_ce_backward(probs, labels, masks, begin, end, grad):
    if begin == end:
        return
    check(lib().gai_softmax_ce_backward_ld(probs.shape[1], begin, end, _p(masks), _p(labels), _f32(probs), probs.stride(0), _f32(grad),
                                           grad.stride(0), end - begin, _stream()), "gai_softmax_ce_backward_ld")


def softmax_ce_backward_scaled(probs, labels, masks, begin, end, grad, denom):
    """grad rows may be wider than ncls (grad.stride(0) is the leading dimension); scaled by 1/denom (the global range length)."""
    check(lib().gai_softmax_ce_backward_ld(probs.shape[1], begin, end, _p(masks), _p(labels), _f32(probs), probs.stride(0), _f32(grad),
                                           grad.stride(0), denom, _stream()), "gai_softmax_ce_backward_ld")


def masked_loss_accuracy(logits, labels, masks, begin, end, losses, stats=None):
    stats = torch.empty(3, dtype=torch.float32, device=logits.device) if stats is None else stats
    check(lib().gai_masked_loss_accuracy_ld(logits.shape[1], begin, end, _p(masks), _p(labels), _f32(logits), logits.stride(0), _f32(losses),
                                            _f32(stats), _stream()), "gai_masked_loss_accuracy_ld")
    return stats


def sigmoid_ce_forward(logits, labels_multi, masks, begin, end, probs, losses):
    """sigmoid_loss_layer::forward: probs = sigmoid(logits), losses[i] = multi-label cross-entropy; labels_multi uint8 [nv, ncls]."""
    check(lib().gai_sigmoid_ce_forward_ld(logits.shape[1], begin, end, _p(masks), _p(labels_multi), _f32(logits), logits.stride(0), _f32(probs),
                                          probs.stride(0), _f32(losses), _stream()), "gai_sigmoid_ce_forward_ld")


def sigmoid_ce_backward(probs, labels_multi, masks, begin, end, grad, denom=None):
    check(lib().gai_sigmoid_ce_backward_ld(probs.shape[1], begin, end, _p(masks), _p(labels_multi), _f32(probs), probs.stride(0), _f32(grad),
                                           grad.stride(0), denom if denom is not None else max(end - begin, 1), _stream()), "gai_sigmoid_ce_backward_ld")


def masked_loss_mean(losses, masks, begin, end):
    stats = torch.empty(3, dtype=torch.float32, device=losses.device)
    check(lib().gai_masked_loss_mean(begin, end, _p(masks), _f32(losses), _f32(stats), _stream()), "gai_masked_loss_mean")
    return stats


def masked_f1_micro(preds, labels_multi, masks, begin, end):
    out = torch.empty(1, dtype=torch.float32, device=preds.device)
    check(lib().gai_masked_f1_micro(preds.shape[1], begin, end, _p(masks), _p(labels_multi), _f32(preds), preds.stride(0), _f32(out), _stream()),
          "gai_masked_f1_micro")
    return out


def adam_update(dW, W, m, v, lr, b1_t, b2_t, b1=0.9, b2=0.999, eps=1e-8):
    check(lib().gai_adam_update(W.numel(), _f32(dW), _f32(W), _f32(m), _f32(v), lr, b1, b2, b1_t, b2_t, eps, _stream()), "gai_adam_update")


def gat_forward(g, z, alpha_l, alpha_r, slope=0.2, flags=0):
    F = z.shape[1]
    temp = torch.empty(max(g.nnz, 1), dtype=torch.float32, device=z.device)
    norm = torch.empty(max(g.nnz, 1), dtype=torch.float32, device=z.device)
    out = torch.empty(g.nv, F, dtype=torch.float32, device=z.device)
    check(lib().gai_gat_forward(g.handle, F, _f32(z), _f32(alpha_l), _f32(alpha_r), slope, _f32(temp), _f32(norm), _f32(out), flags, _stream()),
          "gai_gat_forward")
    return out, temp, norm


def gat_forward_heads(g, z, heads, alpha_l, alpha_r, slope=0.2, flags=0):
    """Multi-head attention forward (extension; heads == 1 is gat_forward). Score arrays are edge-major [nnz x heads]."""
    F = z.shape[1]
    temp = torch.empty(max(g.nnz, 1) * heads, dtype=torch.float32, device=z.device)
    norm = torch.empty(max(g.nnz, 1) * heads, dtype=torch.float32, device=z.device)
    out = torch.empty(g.nv, F, dtype=torch.float32, device=z.device)
    check(lib().gai_gat_forward_heads_ld(g.handle, F, heads, _f32(z), z.stride(0), _f32(alpha_l), _f32(alpha_r), slope, _f32(temp), _f32(norm), _f32(out),
                                         out.stride(0), flags, _stream()), "gai_gat_forward_heads_ld")
    return out, temp, norm


def gat_backward_heads(g, z, heads, grad_in, temp, norm, slope=0.2):
    F = z.shape[1]
    ds = torch.empty(max(g.nnz, 1) * heads, dtype=torch.float32, device=z.device)
    dal = torch.empty(F, dtype=torch.float32, device=z.device)
    dar = torch.empty(F, dtype=torch.float32, device=z.device)
    dz = torch.empty_like(z)
    check(lib().gai_gat_backward_heads_ld(g.handle, F, heads, _f32(z), z.stride(0), _f32(grad_in), grad_in.stride(0), slope, _f32(temp), _f32(norm),
                                          _f32(ds), _f32(dal), _f32(dar), _f32(dz), dz.stride(0), _stream()), "gai_gat_backward_heads_ld")
    return dz, dal, dar, ds


def gat_backward(g, z, grad_in, temp, norm, slope=0.2, dz=None):
    F = z.shape[1]
    ds = torch.empty(max(g.nnz, 1), dtype=torch.float32, device=z.device)
    dal = torch.empty(F, dtype=torch.float32, device=z.device)
    dar = torch.empty(F, dtype=torch.float32, device=z.device)
    dz = torch.empty_like(z) if dz is None else dz
    check(lib().gai_gat_backward(g.handle, F, _f32(z), _f32(grad_in), slope, _f32(temp), _f32(norm), _f32(ds), _f32(dal), _f32(dar), _f32(dz),
                                 _stream()), "gai_gat_backward")
    return dz, dal, dar, ds


def gather_rows(ids, src, out=None):
    F = src.shape[1]
    out = torch.empty(ids.numel(), F, dtype=torch.float32, device=ids.device) if out is None else out
    check(lib().gai_gather_rows(ids.numel(), _p(ids), F, _f32(src), src.stride(0), _f32(out), out.stride(0), _stream()), "gai_gather_rows")
    return out


def partition1d(rowptr64, colidx, nparts, part):
    """Host-side 1D master+halo partition (bit-exact vs the reference partitioner). Returns dict of numpy arrays."""
    rp = np.ascontiguousarray(rowptr64, np.int64)
    ci = np.ascontiguousarray(colidx, np.uint32)
    nv = len(rp) - 1
    m, ne = C.c_int64(), C.c_int64()
    lb, le = C.c_uint32(), C.c_uint32()
    L = lib()
    check(L.gai_partition1d_h(nv, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), nparts, part, None, None, None,
                              C.byref(m), C.byref(ne), C.byref(lb), C.byref(le)), "gai_partition1d_h")
    idx = np.empty(m.value, np.uint32)
    srp = np.empty(m.value + 1, np.int64)
    sci = np.empty(max(ne.value, 1), np.uint32)
    check(L.gai_partition1d_h(nv, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), nparts, part, idx.ctypes.data_as(C.c_void_p),
                              srp.ctypes.data_as(C.c_void_p), sci.ctypes.data_as(C.c_void_p), C.byref(m), C.byref(ne), C.byref(lb), C.byref(le)),
          "gai_partition1d_h")
    return dict(idx_map=idx, rowptr=srp, colidx=sci[: ne.value], local_begin=lb.value, local_end=le.value)
