"""ctypes binding of libdgcnn_b200.so (include/dgcnn_b200.h).  No fallback: a missing library is fatal."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdgcnn_b200.so")

c_void_p, c_int, c_float, c_double, c_size_t, c_int64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_int64


class EdgeConvArgs(C.Structure):
    """mirror of `dgcnn_edgeconv_args`"""
    _fields_ = [("rows", c_void_p), ("ld", c_int), ("idx", c_void_p),
                ("B", c_int), ("C", c_int), ("N", c_int), ("k", c_int),
                ("w_a", c_void_p), ("w_b", c_void_p), ("out", c_void_p), ("ld_out", c_int), ("v_scratch", c_void_p),
                ("impl", c_int)]


_P = c_void_p
_ARGS = C.POINTER(EdgeConvArgs)
# name -> (restype, argtypes); must list every symbol declared in include/dgcnn_b200.h
SIGNATURES = {
    "dgcnn_version": (c_int, []),
    "dgcnn_last_error": (C.c_char_p, []),
    "dgcnn_num_sms": (c_int, []),
    "dgcnn_pack_rows_f32": (c_int, [_P, c_int64, c_int64, c_int64, c_int, c_int, c_int, _P, c_int, _P]),
    "dgcnn_unpack_rows_f32": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "dgcnn_knn_f32": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "dgcnn_knn_tc_supported": (c_int, [c_int, c_int, c_int]),
    "dgcnn_knn_tc_workspace_bytes": (c_size_t, [c_int, c_int]),
    "dgcnn_knn_tc_f32": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "dgcnn_knn_tc_fallback_stats": (c_int, [_P, c_int, c_int, c_int, C.POINTER(c_int), _P]),
    "dgcnn_knn_small_supported": (c_int, [c_int, c_int, c_int]),
    "dgcnn_knn_small_workspace_bytes": (c_size_t, [c_int, c_int]),
    "dgcnn_knn_small_f32": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "dgcnn_graph_feature_f32": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "dgcnn_edgeconv_fwd_eval": (c_int, [_ARGS, _P, _P, _P]),
    "dgcnn_edgeconv_partials_bytes": (c_size_t, []),
    "dgcnn_edgeconv_conv_a_train": (c_int, [_ARGS, _P, _P, _P, _P, _P, _P, _P]),
    "dgcnn_edgeconv_conv_b_train": (c_int, [_ARGS, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dgcnn_bn_finalize": (c_int, [_P, c_int, c_double, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P]),
    "dgcnn_edgeconv_select": (c_int, [c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P]),
    "dgcnn_edgeconv_bwd_prep": (c_int, [c_int, _P, c_int, _P, _P, _P, _P, _P]),
    "dgcnn_bn_bwd_reduce": (c_int, [_P, c_int, c_double, _P, _P, _P, _P, _P, _P, _P]),
    "dgcnn_edgeconv_dw_partials_bytes": (c_size_t, []),
    "dgcnn_edgeconv_bwd_b": (c_int, [_ARGS, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dgcnn_edgeconv_bwd_a": (c_int, [_ARGS, _P, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "dgcnn_edgeconv_lin_extrema": (c_int, [_ARGS, _P, c_int, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "dgcnn_edgeconv_lin_conv_a": (c_int, [_ARGS, _P, c_int, _P, c_int, _P, _P, _P]),
    "dgcnn_edgeconv_lin_conv_a_train": (c_int, [_ARGS, _P, c_int, _P, c_int, _P, _P, _P, _P]),
    "dgcnn_edgeconv_bwd_b_fused": (c_int, [_ARGS, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P]),
    "dgcnn_edgeconv_bwd_a_point_stats": (c_int, [c_int, _P, c_int, _P, c_int, _P, _P, c_int, _P, _P]),
    "dgcnn_edgeconv_bwd_a_point_apply": (c_int, [c_int, c_int, _P, _P, c_int, _P, _P, _P, _P, _P]),
    "dgcnn_edgeconv_bwd_a_lin_edges": (c_int, [_ARGS, _P, c_int, _P, c_int, _P, _P, _P, _P, _P, c_int, _P, _P]),
    "dgcnn_graph_transpose_workspace_bytes": (c_size_t, [c_int, c_int]),
    "dgcnn_graph_transpose": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "dgcnn_gather_in_edges_f32": (c_int, [_P, _P, _P, c_int, _P, c_int, _P]),
    "dgcnn_global_feature_scatter_det_f32": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P]),
    "dgcnn_edgeconv_bwd_a_lin": (c_int, [_ARGS, _P, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "dgcnn_tc_gemm_colstats_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "dgcnn_tc_gemm_bn_lrelu_segmax_f32": (c_int, [_P, c_int64, c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P,
                                                  c_float, c_float, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "dgcnn_global_feature_bwd_sparse_f32": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P, _P]),
    "dgcnn_vecmat_f32": (c_int, [_P, _P, c_int, c_int, c_int, c_float, _P, _P]),
    "dgcnn_global_feature_bwd_coef_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    "dgcnn_global_feature_bwd_dw_f32": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P]),
    "dgcnn_edgeconv_lin_weight_f32": (c_int, [_P, c_int, _P, _P]),
    "dgcnn_edgeconv_lin_weight_bwd_f32": (c_int, [_P, c_int, _P, _P]),
    "dgcnn_rows_bn_supported": (c_int, [c_int]),
    "dgcnn_rows_bn_workspace_bytes": (c_size_t, [c_int]),
    "dgcnn_rows_bn_lrelu_fwd": (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, c_float, c_float, c_int, _P, _P, _P, _P, _P, _P,
                                        _P, c_size_t, _P]),
    "dgcnn_rows_bn_lrelu_bwd": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "dgcnn_linear_narrow_supported": (c_int, [c_int, c_int]),
    "dgcnn_linear_narrow_fwd_f32": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "dgcnn_linear_narrow_bwd_f32": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P]),
    "dgcnn_linear_thin_supported": (c_int, [c_int, c_int]),
    "dgcnn_linear_thin_fwd_f32": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "dgcnn_linear_thin_bwd_f32": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P]),
    "dgcnn_rows_bn_lrelu_fwd_planes": (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, c_size_t,
                                               c_int, _P, c_size_t, _P]),
    "dgcnn_rows_bn_lrelu_bwd_planes": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P, _P, _P, _P, c_size_t, c_int, _P, _P, _P, _P, c_size_t,
                                               _P, c_size_t, _P]),
    "dgcnn_rows_colsum_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dgcnn_rows_colsum_f32": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "dgcnn_seg_loss_workspace_bytes": (c_size_t, [c_int, c_int]),
    "dgcnn_seg_loss_f32": (c_int, [_P, c_int64, c_int64, c_int64, c_int, c_int, c_int, _P, _P, c_float, c_float, _P,
                                   _P, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "dgcnn_seg_predict_f32": (c_int, [_P, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_float, c_int64, _P, _P, _P]),
    "dgcnn_sample_points_workspace_bytes": (c_size_t, [c_int]),
    "dgcnn_sample_points": (c_int, [_P, c_int, c_int, c_int, C.c_uint64, _P, _P, c_size_t, _P]),
    "dgcnn_gather_points": (c_int, [_P, c_int, c_int, _P, _P, c_int, c_int, _P, c_int64, c_int64, _P, _P]),
    "dgcnn_tc_planes_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dgcnn_tc_split_planes_f32": (c_int, [_P, c_int64, c_int64, c_int, c_int, c_int, _P, c_size_t, _P]),
    "dgcnn_tc_gemm_planes_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "dgcnn_tc_gemm_planes_f32": (c_int, [c_int, _P, c_int, c_int, _P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int,
                                         _P, c_size_t, _P]),
    "dgcnn_tc_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "dgcnn_tc_gemm_f32": (c_int, [_P, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int, c_int, c_int, c_int, c_int,
                                  _P, c_size_t, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ahn3-dgcnn.pytorch_b200/csrc`).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DgcnnError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = load().dgcnn_last_error().decode("utf-8", "replace")
        raise DgcnnError(f"libdgcnn_b200 error {rc}: {msg}")
