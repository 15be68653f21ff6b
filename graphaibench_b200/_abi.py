"""ctypes binding of include/gai_b200.h (libgai_b200.so). No fallback: if the library or a CUDA device is missing,
calls fail loudly (GaiError / OSError)."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GAI_B200_LIB", os.path.join(PKG, "libgai_b200.so"))  # override: A/B builds of the same ABI

c_f32p = C.c_void_p  # device pointers travel as integers
c_u32p = C.c_void_p
c_u8p = C.c_void_p
c_stream = C.c_void_p


class GaiError(RuntimeError):
    pass


_SIGS = {
    "gai_last_error": (C.c_char_p, []),
    "gai_version": (C.c_int, []),
    "gai_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gai_set_device": (C.c_int, [C.c_int]),
    "gai_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "gai_free": (C.c_int, [C.c_void_p]),
    "gai_memset": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, c_stream]),
    "gai_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    "gai_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    "gai_memcpy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    "gai_memcpy2d": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, c_stream]),
    "gai_stream_sync": (C.c_int, [c_stream]),
    "gai_stream_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "gai_stream_destroy": (C.c_int, [c_stream]),
    "gai_stream_wait_event": (C.c_int, [c_stream, C.c_void_p]),
    "gai_host_alloc_pinned": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "gai_host_free_pinned": (C.c_int, [C.c_void_p]),
    "gai_event_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "gai_event_record": (C.c_int, [C.c_void_p, c_stream]),
    "gai_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "gai_event_destroy": (C.c_int, [C.c_void_p]),
    "gai_launch_count": (C.c_uint64, []),
    "gai_add_selfloop_h": (C.c_int, [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gai_coo_to_csr": (C.c_int, [C.c_uint32, C.c_uint64, c_u32p, c_u32p, C.c_int, c_stream, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "gai_add_selfloop_d": (C.c_int, [C.c_uint32, C.c_uint32, c_u32p, c_u32p, c_u32p, c_u32p, c_stream]),
    "gai_induced_subgraph": (C.c_int, [C.c_void_p, C.c_uint32, c_u32p, c_stream, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "gai_csr_create": (C.c_int, [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, c_stream, C.POINTER(C.c_void_p)]),
    "gai_csr_create_device": (C.c_int, [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, c_stream, C.POINTER(C.c_void_p)]),
    "gai_csr_destroy": (C.c_int, [C.c_void_p]),
    "gai_csr_nv": (C.c_uint32, [C.c_void_p]),
    "gai_csr_nnz": (C.c_uint64, [C.c_void_p]),
    "gai_csr_rowptr": (C.c_void_p, [C.c_void_p]),
    "gai_csr_colidx": (C.c_void_p, [C.c_void_p]),
    "gai_csr_vertex_norm": (C.c_void_p, [C.c_void_p]),
    "gai_csr_set_norms": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_stream]),
    "gai_csr_num_hub_rows": (C.c_uint32, [C.c_void_p]),
    "gai_csr_max_degree": (C.c_uint32, [C.c_void_p]),
    "gai_triangle_count_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), c_stream]),
    "gai_csr_set_row_segments": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_stream]),
    "gai_csr_build_transpose": (C.c_int, [C.c_void_p, c_stream]),
    "gai_csr_transpose_perm": (C.c_void_p, [C.c_void_p]),
    "gai_spmm_gcn": (C.c_int, [C.c_void_p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, c_f32p, c_stream]),
    "gai_spmm_mean": (C.c_int, [C.c_void_p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p, c_stream]),
    "gai_spmm_gcn_masked": (C.c_int, [C.c_void_p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, c_f32p, C.c_void_p, C.c_int, c_stream]),
    "gai_spmm_mean_masked": (C.c_int, [C.c_void_p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p, C.c_int, c_stream]),
    "gai_spmm_edge": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_u32p, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, c_f32p, c_stream]),
    "gai_spmm_gcn_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, c_f32p, c_stream]),
    "gai_spmm_mean_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p, c_stream]),
    "gai_gat_forward": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, c_f32p, C.c_float, c_f32p, c_f32p, c_f32p, C.c_int, c_stream]),
    "gai_gat_backward": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, C.c_float, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_stream]),
    "gai_spmm_edge_heads": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_f32p, c_u32p, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, c_f32p, c_stream]),
    "gai_gat_forward_heads_ld": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_f32p, C.c_size_t, c_f32p, c_f32p, C.c_float, c_f32p, c_f32p, c_f32p, C.c_size_t,
                                           C.c_int, c_stream]),
    "gai_gat_backward_heads_ld": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_float, c_f32p, c_f32p, c_f32p, c_f32p,
                                            c_f32p, c_f32p, C.c_size_t, c_stream]),
    "gai_gat_forward_ld": (C.c_int, [C.c_void_p, C.c_int, c_f32p, C.c_size_t, c_f32p, c_f32p, C.c_float, c_f32p, c_f32p, c_f32p, C.c_size_t, C.c_int, c_stream]),
    "gai_gat_backward_ld": (C.c_int, [C.c_void_p, C.c_int, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_float, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                                      c_f32p, C.c_size_t, c_stream]),
    "gai_matmul": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "gai_matmul_ld": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t,
                                C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "gai_matmul_kcat": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t,
                                  c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, c_stream]),
    "gai_matmul_relu_bits": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_int,
                                       C.c_void_p, C.c_size_t, c_stream]),
    "gai_matmul_mask": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_int,
                                  C.c_void_p, C.c_size_t, C.c_int, c_stream]),
    "gai_matmul_ncat": (C.c_int, [C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_size_t,
                                  c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_int, c_stream]),
    "gai_wgrad_two_a": (C.c_int, [C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_size_t,
                                  c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_stream]),
    "gai_wgrad_two_b": (C.c_int, [C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_size_t,
                                  c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_stream]),
    "gai_d_relu_ld": (C.c_int, [C.c_size_t, C.c_int, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_stream]),
    "gai_set_gemm_mode": (C.c_int, [C.c_int]),
    "gai_get_gemm_mode": (C.c_int, []),
    "gai_relu": (C.c_int, [C.c_size_t, c_f32p, c_f32p, c_stream]),
    "gai_d_relu": (C.c_int, [C.c_size_t, c_f32p, c_f32p, c_f32p, c_stream]),
    "gai_dropout": (C.c_int, [C.c_size_t, C.c_float, C.c_float, C.c_uint64, C.c_uint64, c_f32p, c_u8p, c_f32p, c_stream]),
    "gai_d_dropout": (C.c_int, [C.c_size_t, C.c_float, c_f32p, c_u8p, c_f32p, c_stream]),
    "gai_fill": (C.c_int, [C.c_size_t, C.c_float, c_f32p, c_stream]),
    "gai_l2norm": (C.c_int, [C.c_int, C.c_int, c_f32p, c_f32p, c_stream]),
    "gai_d_l2norm": (C.c_int, [C.c_int, C.c_int, c_f32p, c_f32p, c_f32p, c_stream]),
    "gai_l2norm_ld": (C.c_int, [C.c_int, C.c_int, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_stream]),
    "gai_d_l2norm_ld": (C.c_int, [C.c_int, C.c_int, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_stream]),
    "gai_softmax_ce_forward_ld": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_f32p, c_stream]),
    "gai_softmax_ce_forward_stats_ld": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_f32p,
                                                  c_f32p, c_stream]),
    "gai_softmax_ce_backward_ld": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_uint64, c_stream]),
    "gai_masked_loss_accuracy_ld": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, C.c_size_t, c_f32p, c_f32p, c_stream]),
    "gai_softmax_ce_forward": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, c_f32p, c_f32p, c_stream]),
    "gai_softmax_ce_backward": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, c_f32p, c_stream]),
    "gai_softmax_ce_backward_scaled": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, c_f32p, C.c_int, C.c_uint64, c_stream]),
    "gai_masked_loss_accuracy": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, c_f32p, c_f32p, c_stream]),
    "gai_sigmoid_ce_forward_ld": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, C.c_size_t, c_f32p, C.c_size_t, c_f32p, c_stream]),
    "gai_sigmoid_ce_backward_ld": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, C.c_size_t, c_f32p, C.c_size_t, C.c_uint64, c_stream]),
    "gai_masked_loss_mean": (C.c_int, [C.c_size_t, C.c_size_t, c_u8p, c_f32p, c_f32p, c_stream]),
    "gai_masked_f1_micro": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, c_u8p, c_u8p, c_f32p, C.c_size_t, c_f32p, c_stream]),
    "gai_adam_update": (C.c_int, [C.c_size_t, c_f32p, c_f32p, c_f32p, c_f32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_float, c_stream]),
    "gai_partition1d_h": (C.c_int, [C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "gai_peers_create": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_stream, C.POINTER(C.c_void_p)]),
    "gai_peers_destroy": (C.c_int, [C.c_void_p]),
    "gai_peers_rank": (C.c_int, [C.c_void_p]),
    "gai_peers_world": (C.c_int, [C.c_void_p]),
    "gai_peers_register": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "gai_peers_barrier": (C.c_int, [C.c_void_p, c_stream]),
    "gai_peers_barrier_on": (C.c_int, [C.c_void_p, C.c_int, c_stream]),
    "gai_peers_error": (C.c_int, [C.c_void_p, c_stream]),
    "gai_halo_plan_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, c_stream, C.POINTER(C.c_void_p)]),
    "gai_halo_plan_destroy": (C.c_int, [C.c_void_p]),
    "gai_halo_pull": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, c_f32p, C.c_size_t, C.c_int, c_stream]),
    "gai_halo_pull_cols": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, c_f32p, C.c_size_t, C.c_int, C.c_int, c_stream]),
    "gai_peers_combine": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_int, c_f32p, c_stream]),
    "gai_spmm_rows_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, c_f32p, c_u32p, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int,
                                   c_f32p, C.c_void_p, C.c_int, c_f32p, C.c_uint32, c_stream]),
    "gai_csr_mean_norm": (C.c_void_p, [C.c_void_p]),
    "gai_gather_rows": (C.c_int, [C.c_size_t, c_u32p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, c_stream]),
}

# every symbol include/gai_b200.h declares (tests/test_abi.py parses the header and compares)
EXPORTED = tuple(_SIGS)

_lib = None


def lib():
    """Load libgai_b200.so (built by graphaibench_b200/build.py). Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not built: run `python -m graphaibench_b200.build` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().gai_last_error()
        raise GaiError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
