"""ctypes binding of libpagegeom.so (include/pagegeom.h).

The library is the product: if it is missing this module raises — there is no CPU or
PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpagegeom.so")

PG_OK = 0
PG_ERR_WORKSPACE = 3
PG_FLAG_PLAIN_TEXT = 1
PG_FLAG_TITLE = 2
PG_WIDTH_HIST_BINS = 16384
PG_COL_HIST_BINS = 1001
PG_COL_SPAN_BYTES = 32
PG_NMS_CLASS_AGNOSTIC = 1
PG_NMS_FP32 = 2
PG_JSON_INT_LITERALS_TO_HOST = 1


class PgTileInfo(C.Structure):
    _fields_ = [
        ("x_start", C.c_double), ("y_start", C.c_double), ("x_end", C.c_double), ("y_end", C.c_double),
        ("grid_rows", C.c_int32), ("grid_cols", C.c_int32), ("row", C.c_int32), ("col", C.c_int32),
        ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
        ("new_w", C.c_int32), ("new_h", C.c_int32), ("pad_l", C.c_int32), ("pad_t", C.c_int32),
        ("out_w", C.c_int32), ("out_h", C.c_int32), ("out_offset", C.c_int64),
    ]


class PageGeomError(RuntimeError):
    pass


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F64 = C.c_double

# name -> (restype, argtypes); every symbol declared in include/pagegeom.h
SIGNATURES = {
    "pg_last_error": (C.c_char_p, []),
    "pg_version": (C.c_int, []),
    "pg_build_checked": (C.c_int, []),
    "pg_pinned_alloc": (C.c_int, [C.c_size_t, C.c_int32, C.POINTER(C.c_void_p)]),
    "pg_pinned_free": (C.c_int, [C.c_void_p]),
    "pg_device_info": (C.c_int, [_P, _P, _P]),
    "pg_tile_plan_create": (C.c_int, [_I32, _I32, _P, _P, _I32, _F64, _I32, _I32, _I32, _I32, _P]),
    "pg_tile_plan_create_ex": (C.c_int, [_I32, _I32, _I32, _P, _P, _I32, _F64, _I32, _I32, _I32, _I32, _P]),
    "pg_tile_plan_destroy": (None, [_P]),
    "pg_tile_plan_num_tiles": (_I32, [_P]),
    "pg_tile_plan_tile": (C.c_int, [_P, _I32, _P]),
    "pg_tile_plan_out_elems": (_I64, [_P]),
    "pg_tile_plan_algorithmic_bytes": (_I64, [_P]),
    "pg_tile_letterbox": (C.c_int, [_P, _P, _I32, _I64, _I64, _P, _I64, _P]),
    "pg_tile_letterbox_direct": (C.c_int, [_P, _P, _I32, _I64, _I64, _P, _I64, _P]),
    "pg_tile_batch_create": (C.c_int, [_P, _I32, _P, _I32, _P]),
    "pg_tile_batch_destroy": (None, [_P]),
    "pg_tile_batch_algorithmic_bytes": (_I64, [_P]),
    "pg_tile_batch_bind": (C.c_int, [_P, _P, _P, _P, _P]),
    "pg_tile_letterbox_batch": (C.c_int, [_P, _P]),
    "pg_synth_pages": (C.c_int, [_P, _I32, _I32, _I32, _I64, _I64, C.c_uint64, _I64, _P]),
    "pg_edge_filter": (C.c_int, [_P, _I32, _P, _P, _P, _P, _I32, _F64, _P, _P, _P, _P, _P]),
    "pg_nms_workspace_bytes": (C.c_size_t, [_I64, _I32, _I32]),
    "pg_nms_merge": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I64, _I32, _F64, _P, _P, _P, C.c_size_t, _P]),
    "pg_nms_merge_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I64, _I32, _F64, _I32, _P, _P, _P, C.c_size_t, _P]),
    "pg_nms_stats": (C.c_int, [_P, _P]),
    "pg_class_flags": (C.c_int, [_P, _I64, _F64, _F64, _P, _P]),
    "pg_width_median": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I64, _P, _F64, _P, _P, _P, _P, _P, _P]),
    "pg_column_peaks": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _I32, _F64, _I32,
                                  _P, _P, _P, _P, _I32, _P, _P, _P, _P]),
    "pg_assign_columns": (C.c_int, [_P, _P, _P, _P, _I32, _I64, _P, _P, _I32, _P, _P]),
    "pg_comm_nccl_version": (C.c_int, []),
    "pg_comm_unique_id": (C.c_int, [_P]),
    "pg_comm_create": (C.c_int, [_P, _I32, _I32, _P]),
    "pg_comm_destroy": (None, [_P]),
    "pg_comm_nccl": (_P, [_P]),
    "pg_hist_allreduce": (C.c_int, [_P, C.c_size_t, _P, _P]),
    "pg_jpeg_decoder_create": (C.c_int, [_P]),
    "pg_jpeg_decoder_destroy": (None, [_P]),
    "pg_jpeg_decoder_configure": (C.c_int, [_P, _I32, _I32]),
    "pg_jpeg_decoder_set_files": (C.c_int, [_P, _P, _P, _I32]),
    "pg_jpeg_decoder_image_info": (C.c_int, [_P, _I32, _P, _P, _P]),
    "pg_jpeg_workspace_bytes": (_I64, [_P]),
    "pg_jpeg_decode": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P]),
    "pg_jpeg_stage_tables": (C.c_int, [_P, _P, _P, _P, _I64, _P]),
    "pg_jpeg_decode_status": (C.c_int, [_P, _P]),
    "pg_hostcheck_jpeg_decode": (C.c_int, [_P, _I64, _I32, _I32, _P, _I64, _P, _P, _P]),
    "pg_json_workspace_bytes": (_I64, [_I64, _I32]),
    "pg_json_combined": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I32, _I64, _I32, _P, _P, _P, _P, _P, _I64, _P,
                                   _P, _I64, _P]),
    "pg_json_segments_workspace_bytes": (_I64, [_I64, _I32]),
    "pg_json_segments": (C.c_int, [_P, _I32, _P, _I64, _P, _I32, _P, _P, _P, _P, _I64, _P, _P, _I64, _P]),
    "pg_json_parse_block_bytes": (_I32, []),
    "pg_json_parse_workspace_bytes": (_I64, [_I64]),
    "pg_json_parse_numbers": (C.c_int, [_P, _P, _I32, _P, _I64, _P, _I64, _P, _P, _I32, _P, _I64, _P]),
    "pg_hostcheck_format_double": (_I32, [_F64, _P]),
    "pg_hostcheck_format_doubles": (_I64, [_P, _I64, _P, _P]),
    "pg_hostcheck_parse_numbers": (_I64, [_P, _I64, _P, _I64, _P, _P]),
    "pg_hostcheck_iou": (_F64, [_P, _P]),
    "pg_hostcheck_iou_gt": (_I32, [_P, _P, _F64]),
    "pg_hostcheck_iou_gt_f32": (_I32, [_P, _P, _F64]),
    "pg_hostcheck_edge_touch": (_I32, [_P, _P, _I32, _I32, _F64]),
    "pg_hostcheck_density_weight": (_F64, [_I32, _I32, _I32, _I32]),
    "pg_hostcheck_density_rcp_mismatches": (_I64, [_I32, _I32]),
    "pg_hostcheck_resize_row": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libpagegeom.so; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("PAGEGEOM_LIB") or LIB_PATH  # A/B builds of the same sources (scripts/gpu_*.sh); default: the in-tree build
        if not os.path.exists(path):
            raise PageGeomError(
                f"{path} not found: build it with `python -m multimodal_embeddings_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != PG_OK:
        msg = lib().pg_last_error()
        raise PageGeomError(f"libpagegeom error {rc}: {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
