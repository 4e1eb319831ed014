"""ctypes binding of libbgnn.so (the C ABI declared in include/bgnn.h).

There is no CPU fallback: if the shared library is missing this module raises at
import time, and every compute entry point raises when CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BGNN_LIB") or os.path.join(HERE, "libbgnn.so")      # BGNN_LIB: instrumented dev builds (tools/)

MAX_OFFSETS, MAX_NODE_FEATS, MAX_EDGE_FEATS = 24, 8, 4
E_INVALID, E_CUDA, E_UNSUPPORTED, E_WORKSPACE = -1, -2, -3, -4
PREC_F32, PREC_BF16 = 0, 1
AGG_GCN, AGG_MEAN, AGG_GIN = 0, 1, 2
OPTIONS = dict(tensor_cores=0, presplit=1, window=2, tile5=3, tile5_window=4, ragged_patches=5)

NF_CODES = dict(depth=0, local_mean=1, local_std=2, gradient_x=3, gradient_y=4,
                gradient_magnitude=5, curvature=6, uncertainty=7)
EF_CODES = dict(distance=0, depth_difference=1, slope=2)
EF_ZERO = 3


class Cells(C.Structure):
    _fields_ = [("n_grids", C.c_int32), ("n_cells", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32),
                ("row_stride", C.c_int64), ("rx", C.c_double), ("ry", C.c_double),
                ("cell_off", C.c_void_p), ("grid_rows", C.c_void_p), ("grid_cols", C.c_void_p),
                ("grid_res", C.c_void_p), ("depth", C.c_void_p), ("mask", C.c_void_p), ("unc", C.c_void_p),
                ("use_nodata", C.c_int32), ("nodata", C.c_float)]


class GraphCfg(C.Structure):
    _fields_ = [("n_offsets", C.c_int32), ("off_dr", C.c_int32 * MAX_OFFSETS), ("off_dc", C.c_int32 * MAX_OFFSETS),
                ("n_node_feats", C.c_int32), ("node_feats", C.c_int32 * MAX_NODE_FEATS),
                ("n_edge_feats", C.c_int32), ("edge_feats", C.c_int32 * MAX_EDGE_FEATS)]


class ModelDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("num_classes", C.c_int32), ("predict_correction", C.c_int32), ("edge_dim", C.c_int32),
                ("bn_eps", C.c_float), ("negative_slope", C.c_float)]


class GridTopo(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("node_id", C.c_void_p), ("n_offsets", C.c_int32),
                ("off_dr", C.c_int32 * MAX_OFFSETS), ("off_dc", C.c_int32 * MAX_OFFSETS), ("grid_eattr", C.c_void_p),
                ("patches", C.c_void_p), ("n_patches", C.c_int32)]


class Survey(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int64), ("cont", C.c_void_p * 4), ("wsum", C.c_void_p * 4),
                ("n_cont", C.c_int32), ("discrete", C.c_void_p), ("tracker", C.c_void_p), ("f64_mask", C.c_int32)]


class Sidecar(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int64), ("classification", C.c_void_p), ("confidence", C.c_void_p),
                ("correction", C.c_void_p), ("valid_mask", C.c_void_p), ("owner", C.c_void_p)]


SIDECAR_DESC = 6

# every symbol include/bgnn.h declares: name -> (restype, argtypes)
P, I32, I64, F32, SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
SYMBOLS = {
    "bgnn_version": (C.c_int, []),
    "bgnn_last_error": (C.c_char_p, []),
    "bgnn_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "bgnn_graph_scratch_bytes": (SZ, [I64, I32]),
    "bgnn_graph_count": (C.c_int, [C.POINTER(Cells), C.POINTER(GraphCfg), P, SZ, C.POINTER(I64), P]),
    "bgnn_graph_emit": (C.c_int, [C.POINTER(Cells), C.POINTER(GraphCfg), P, SZ, P, P, P, P, P, P, P, P, P]),
    "bgnn_graph_emit_bands": (C.c_int, [C.POINTER(Cells), C.POINTER(GraphCfg), P, SZ, P, I32, P, P, P, P, P, P, P]),
    "bgnn_graph_coo": (C.c_int, [C.POINTER(Cells), C.POINTER(GraphCfg), P, SZ, P, I64, P, P, P]),
    "bgnn_model_param_floats": (I64, [C.POINTER(ModelDesc)]),
    "bgnn_model_create": (C.c_int, [C.POINTER(ModelDesc), P, I64, C.c_int, C.POINTER(P)]),
    "bgnn_model_destroy": (None, [P]),
    "bgnn_gat_workspace_bytes": (SZ, [P, I64]),
    "bgnn_model_set_precision": (C.c_int, [P, C.c_int]),
    "bgnn_model_precision": (C.c_int, [P]),
    "bgnn_test_proj_bf16": (C.c_int, [P, I64, C.c_int, P, C.c_int, P, P, P, P, C.c_int, P]),
    "bgnn_gat_forward": (C.c_int, [P, I64, I64, P, P, P, P, P, SZ, P, P, P, P, P, P, F32, F32, P, P, P]),
    "bgnn_gat_grid_path": (C.c_int, [P, C.POINTER(GridTopo), I64]),
    "bgnn_csr_aggregate": (C.c_int, [C.c_int, I64, C.c_int, P, I64, P, P, P, C.c_int, F32, P, P, P, C.c_int, P, I64, C.c_int, P]),
    "bgnn_linear": (C.c_int, [P, I64, I64, C.c_int, P, P, C.c_int, C.c_int, P, I64, P]),
    "bgnn_heads_finish": (C.c_int, [P, I64, C.c_int, C.c_int, C.c_int, P, P, P, P, P, P, P, P, P, P, P, P, F32, F32, P]),
    "bgnn_scatter_tile": (C.c_int, [P, I64, P, P, P, P, F32, P, P, P, P]),
    "bgnn_graph_to_grid": (C.c_int, [P, I64, P, F32, P, P]),
    "bgnn_stitch_init": (C.c_int, [C.POINTER(Survey), P]),
    "bgnn_stitch_add_tile": (C.c_int, [C.POINTER(Survey), I64, I64, I32, I32, P, P, C.POINTER(P), P, P, P]),
    "bgnn_stitch_add_tile_nodes": (C.c_int, [C.POINTER(Survey), I64, I64, I32, I32, P, P, P, P, P, P, P, F32, P]),
    "bgnn_stitch_blend_tile": (C.c_int, [P, I64, I64, I64, I64, I32, I32, P, P, P, P]),
    "bgnn_stitch_finalize": (C.c_int, [C.POINTER(Survey), P]),
    "bgnn_apply_corrections": (C.c_int, [P, P, I64, P, P, P, P, F32, P]),
    "bgnn_survey_emit": (C.c_int, [C.POINTER(Survey), P, I32, F32, F32, P, P, I64, P]),
    "bgnn_memcpy2d_async": (C.c_int, [P, SZ, P, SZ, SZ, SZ, C.c_int, P]),
    "bgnn_vr_apply_results": (C.c_int, [P, P, I64, F32, P, P, P, F32, P, P, P, P, P]),
    "bgnn_sidecar_init": (C.c_int, [C.POINTER(Sidecar), P]),
    "bgnn_sidecar_add": (C.c_int, [C.POINTER(Sidecar), I32, P, P, I64, P, P, P, P, F32, P]),
    "bgnn_model_set_option": (C.c_int, [P, C.c_int, C.c_int]),
    "bgnn_test_proj_umma_presplit": (C.c_int, [P, I64, C.c_int, P, C.c_int, P, P, P, P, C.c_int, P]),
    "bgnn_test_proj_umma": (C.c_int, [P, I64, C.c_int, P, C.c_int, P, P, P, P, C.c_int, P]),
    "bgnn_profile_enable": (C.c_int, [C.c_int]),
    "bgnn_profile_collect": (C.c_int, [C.c_int, P, P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python bathymetric-gnn_b200/build.py` "
            "(there is no CPU fallback for the Bathymetric-GNN hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class BgnnError(RuntimeError):
    pass


def check(rc: int):
    """Map C status codes to the reference's Python exception behaviour
    (SURVEY.md 8(b)): invalid argument -> ValueError, everything else -> RuntimeError."""
    if rc == 0:
        return
    msg = (lib.bgnn_last_error() or b"").decode("utf-8", "replace")
    if rc == E_INVALID:
        raise ValueError(msg)
    raise BgnnError(f"libbgnn error {rc}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bathymetric_gnn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t):
    """device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
