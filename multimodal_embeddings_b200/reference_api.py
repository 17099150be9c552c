"""Host-side mirror of the reference's Python functions for the page-geometry path.

Same names, argument meaning, return shapes and error behaviour as the numbered scripts of
calhounpaul/multimodal_embeddings, so the reference's call sites (and the parity tests) can
switch by changing an import.  Each function moves its inputs to the GPU, runs the
corresponding libpagegeom.so kernel(s) and brings the result back; nothing is computed on the
CPU (no fallback: without the library or a CUDA device these raise).
"""
from __future__ import annotations

import logging
import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import PG_FLAG_PLAIN_TEXT, PG_FLAG_TITLE

logger = logging.getLogger("PageGeomB200")


# ------------------------------------------------------------------ stage 1 (1_doclayout_bboxes.py)
def parse_grid_configs(grid_str):
    """1_doclayout_bboxes.py:656-680 (pure string parsing; host logic)."""
    grid_configs = []
    try:
        if grid_str:
            for config in grid_str.split(","):
                config = config.strip()
                if "x" in config:
                    rows, cols = config.split("x")
                    grid_configs.append((int(rows), int(cols)))
    except ValueError as e:
        logger.error(f"Error parsing grid configuration: {str(e)}")
    return grid_configs


def split_image_into_grid(image, rows, cols, overlap_percentage, imgsz=1024, stride=32, auto=True):
    """split_image_into_grid (1_doclayout_bboxes.py:366-444) fused with the detector's
    letterbox/normalise front half (:191-210).

    ``image`` is a path (decoded on the host with cv2.imread like the reference, :381) or an
    already decoded BGR uint8 array.  Returns the reference's list of cell dicts —
    'coordinates' (un-truncated, ints where the clamp hit), 'row', 'col' (1-indexed) — where
    'image' is replaced by 'tensor': the detector-ready fp16 CHW RGB tile on the GPU.
    Unreadable image -> [] (as :382-384)."""
    if isinstance(image, (str, os.PathLike)):
        import cv2
        arr = cv2.imread(str(image))
        if arr is None:
            logger.error(f"Failed to load image for grid splitting: {image}")
            return []
    else:
        arr = np.asarray(image)
    h, w = arr.shape[:2]
    plan = ops.TilePlan(w, h, [(rows, cols)], overlap_percentage, imgsz, stride, auto)
    pages = ops.upload_pages([arr], plan)
    out = plan.run(pages)
    cells = []
    for t, info in enumerate(plan.tiles):
        cells.append({
            "tensor": plan.tile_view(out, 0, t),
            "coordinates": plan.cell_coordinates(t),
            "row": info["row"],
            "col": info["col"],
            "slice": (info["x0"], info["y0"], info["x1"], info["y1"]),
            "letterbox": {k: info[k] for k in ("new_w", "new_h", "pad_l", "pad_t", "out_w", "out_h")},
        })
    return cells


def nms(boxes, scores, iou_threshold: float):
    """Drop-in for ``torchvision.ops.nms(boxes, scores, iou_threshold)`` as called per tile at
    1_doclayout_bboxes.py:219-223: float32 boxes [N,4] / scores [N] (tensor or array) -> int64 tensor of the
    kept indices in decreasing score order.  Class-agnostic greedy NMS evaluated in float32 like
    torchvision's kernel (pg_nms_merge_ex, PG_NMS_CLASS_AGNOSTIC | PG_NMS_FP32)."""
    return nms_per_tile([boxes], [scores], iou_threshold)[0]


def nms_per_tile(boxes_list, scores_list, iou_threshold: float):
    """The per-tile NMS of 1_doclayout_bboxes.py:217-225 for many tiles in ONE launch."""
    from ._lib import PG_NMS_CLASS_AGNOSTIC, PG_NMS_FP32

    def f32(x, shape):
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        return np.asarray(x, np.float32).reshape(shape)

    bs = [f32(b, (-1, 4)) for b in boxes_list]
    ss = [f32(s, (-1,)) for s in scores_list]
    counts = [len(s) for s in ss]
    n = sum(counts)
    if n == 0:
        return [torch.zeros(0, dtype=torch.int64) for _ in bs]
    off = np.concatenate([[0], np.cumsum(counts)])
    kept, n_kept, _ = ops.nms_merge(np.concatenate(bs).astype(np.float64), np.concatenate(ss).astype(np.float64), None,
                                    off, iou_threshold, max_boxes_per_page=max(counts),
                                    mode=PG_NMS_CLASS_AGNOSTIC | PG_NMS_FP32, check_status=True)
    kept, n_kept = kept.cpu().numpy(), n_kept.cpu().numpy()
    return [torch.from_numpy((kept[off[i]: off[i] + n_kept[i]] - off[i]).astype(np.int64)) for i in range(len(bs))]


def translate_coordinates_to_original(boxes, cell_coordinates):
    """1_doclayout_bboxes.py:484-511 on the GPU (pg_edge_filter's translation stage with the
    filter disabled by an infinite page)."""
    if not len(boxes):
        return []
    n = len(boxes)
    cell = [[cell_coordinates["x_start"], cell_coordinates["y_start"], 0.0, 0.0]]
    bp, _, _, _ = ops.edge_filter(np.asarray(boxes, np.float64), np.zeros(n, np.int32), np.asarray(cell, np.float64),
                                  [[0, 0]], [0, n], threshold=0.0, boxes_are_local=True)
    return bp.cpu().numpy().tolist()


# ------------------------------------------------------------------ stage 2 (2_edge_box_filter.py)
def _cell_tuple(cell_coordinates, image_width, image_height):
    if isinstance(cell_coordinates, dict):  # 2_edge_box_filter.py:62-68
        return [cell_coordinates.get("x_start", 0), cell_coordinates.get("y_start", 0),
                cell_coordinates.get("x_end", image_width), cell_coordinates.get("y_end", image_height)]
    return list(cell_coordinates)


def is_box_touching_internal_edge(box, cell_coordinates, image_width, image_height, threshold=10):
    """2_edge_box_filter.py:44-90 for a single box (one-element launch)."""
    _, keep, _, _ = ops.edge_filter([list(box)], [0], [_cell_tuple(cell_coordinates, image_width, image_height)],
                                    [[image_width, image_height]], [0, 1], threshold, boxes_are_local=False,
                                    want_boxes_page=False)
    return not bool(keep[0].item())


def _image_size(path):
    from PIL import Image
    with Image.open(path) as im:  # header only; the reference decodes the whole page (2:194-197)
        return im.width, im.height


def filter_grid_info(grid_info, threshold=10, image_size=None):
    """filter_grid_info (2_edge_box_filter.py:148-237).  Tests every cell's ``boxes_original``
    against its ``cell_coordinates`` in one launch; returns a new grid-info dict with the
    surviving boxes (same key order), or None when the page size cannot be determined
    (:199-203).  ``image_size=(W,H)`` skips the image probe."""
    filtered = {"original_image_path": grid_info["original_image_path"], "cells": []}
    if "grid_config" in grid_info:
        filtered["grid_config"] = grid_info["grid_config"]
    if image_size is None:
        path = grid_info["image_path"] if ("image_path" in grid_info and os.path.exists(grid_info["image_path"])) \
            else grid_info["original_image_path"]
        if not os.path.exists(path):
            logger.warning(f"Original image not found: {path}")
            return None
        try:
            image_size = _image_size(path)
        except Exception:
            logger.warning(f"Could not read original image: {path}")
            return None
    w, h = image_size
    cells = grid_info["cells"]
    counts = [len(c["regions"]["boxes_original"]) for c in cells]
    n = sum(counts)
    kept_lists: List[List[int]] = [[] for _ in cells]
    if n:
        boxes = np.concatenate([np.asarray(c["regions"]["boxes_original"], np.float64).reshape(-1, 4) for c in cells])
        box_cell = np.repeat(np.arange(len(cells), dtype=np.int32), counts)
        cell_arr = np.asarray([_cell_tuple(c["cell_coordinates"], w, h) for c in cells], np.float64)
        _, keep, _, _ = ops.edge_filter(boxes, box_cell, cell_arr, [[w, h]], [0, n], threshold, boxes_are_local=False,
                                        want_boxes_page=False)
        keep = keep.cpu().numpy().astype(bool)
        start = 0
        for ci, cnt in enumerate(counts):
            kept_lists[ci] = np.nonzero(keep[start:start + cnt])[0].tolist()
            start += cnt
    for cell, idx in zip(cells, kept_lists):
        reg = cell["regions"]
        filtered["cells"].append({
            "cell_path": cell["cell_path"], "cell_json_path": cell["cell_json_path"],
            "cell_coordinates": cell["cell_coordinates"], "row": cell.get("row", 0), "col": cell.get("col", 0),
            "regions": {k: [reg[k][i] for i in idx]
                        for k in ("boxes", "boxes_original", "classes", "scores", "class_names")},
        })
    return filtered


def filter_edge_boxes(regions, threshold=10):
    """filter_edge_boxes (2_edge_box_filter.py:92-146), including its coordinate-frame quirk:
    cell-local ``boxes`` are tested against page-coordinate ``cell_coordinates`` with the
    *cell* image size (:110-119)."""
    if "cell_coordinates" not in regions:
        logger.info("Non-grid image detected, not filtering any boxes")
        return regions
    w, h = regions["image_size"]["width"], regions["image_size"]["height"]
    n = len(regions["boxes"])
    idx: List[int] = []
    if n:
        _, keep, _, _ = ops.edge_filter(np.asarray(regions["boxes"], np.float64), np.zeros(n, np.int32),
                                        [_cell_tuple(regions["cell_coordinates"], w, h)], [[w, h]], [0, n],
                                        threshold, boxes_are_local=False, want_boxes_page=False)
        idx = np.nonzero(keep.cpu().numpy())[0].tolist()
    out = {"image_path": regions["image_path"], "image_size": regions["image_size"],
           "parameters": regions["parameters"]}
    for k in ("boxes", "classes", "scores", "class_names"):
        out[k] = [regions[k][i] for i in idx]
    if "boxes_original" in regions:
        out["boxes_original"] = [regions["boxes_original"][i] for i in idx]
    for k in ("cell_coordinates", "original_image_path", "grid_info"):
        if k in regions:
            out[k] = regions[k]
    return out


# ------------------------------------------------------------------ stage 3 (3_combine_grids.py)
def nms_keep_indices(boxes, scores, classes, iou_threshold=0.5) -> List[int]:
    """Positions kept by apply_non_max_suppression, in pick order."""
    n = len(boxes)
    if n == 0:
        return []
    args = (np.asarray(boxes, np.float64), np.asarray(scores, np.float64), np.asarray(classes, np.float64), [0, n],
            iou_threshold)
    kept, n_kept, _ = ops.nms_merge(*args, max_boxes_per_page=n, check_status=True)  # raises on a device error
    k = int(n_kept[0].item())
    return kept[:k].cpu().numpy().tolist()


def calculate_iou(box1, box2):
    """3_combine_grids.py:46-78 evaluated by the kernel's own inline function (host build)."""
    from ._lib import lib, ptr
    a = np.asarray(box1, np.float64)
    b = np.asarray(box2, np.float64)
    return float(lib().pg_hostcheck_iou(ptr(a), ptr(b)))


def apply_non_max_suppression(boxes, scores, classes, class_names, iou_threshold=0.5):
    """apply_non_max_suppression (3_combine_grids.py:80-138): returns
    (filtered_boxes, filtered_scores, filtered_classes, filtered_class_names) in pick order;
    the input lists are not modified; boxes pass through untouched."""
    if not boxes:
        return [], [], [], []
    idx = nms_keep_indices(boxes, scores, classes, iou_threshold)
    return ([boxes[i] for i in idx], [scores[i] for i in idx], [classes[i] for i in idx],
            [class_names[i] for i in idx])


# ------------------------------------------------------------------ stage 4 (4_extract_median_widths.py)
def _flags_from_names(class_names: Sequence[str]) -> np.ndarray:
    return np.asarray([(PG_FLAG_PLAIN_TEXT if n == "plain_text" else 0) | (PG_FLAG_TITLE if n == "title" else 0)
                       for n in class_names], np.uint8)


def median_plain_text_width(boxes, class_names, page_width, min_margin_percent=0.2) -> Tuple[float, int]:
    """bin_widths + calculate_median_width (4_extract_median_widths.py:49-101) over the
    plain_text boxes (:135-141).  Returns (median, number_of_bins); (0, 0) when empty."""
    n = min(len(boxes), len(class_names))
    if n == 0:
        return 0, 0
    med, nb = ops.width_median(np.asarray(boxes[:n], np.float64), _flags_from_names(class_names[:n]), [0, n],
                               [[page_width, 0]], min_margin_percent)
    nb = int(nb[0].item())
    return (np.float64(med[0].item()) if nb else 0), nb


# ------------------------------------------------------------------ stage 5 (5_detect_column_centers.py)
def find_column_centers(boxes, class_names, scores, page_width, page_height, median_width, min_confidence=0.3,
                        verbose=False):
    """find_column_centers (5_detect_column_centers.py:91-224): (column_centers, column_widths)."""
    n = len(boxes)
    if n == 0 or not (median_width > 0):
        if n and verbose:
            logger.warning("median width not positive")
        return [], []
    centers, widths, n_cols = ops.column_peaks(np.asarray(boxes, np.float64), _flags_from_names(class_names),
                                               np.asarray(scores, np.float64), [0, n],
                                               [[page_width, max(1, page_height)]], [float(median_width)],
                                               min_confidence, max_cols=256)
    k = int(n_cols[0].item())
    if k < 0:
        raise RuntimeError("pg_column_peaks: page shape outside kernel limits")
    if k > 256:
        raise RuntimeError(f"pg_column_peaks: {k} columns exceed max_cols")
    c = centers[0, :k].cpu().numpy().tolist()
    w = widths[0, :k].cpu().numpy().tolist()
    if k == 0:
        return [], []
    # the reference returns ints for walked widths and the float median for substituted ones
    w = [int(x) if float(x).is_integer() and x != median_width and x != 2.0 * median_width else x for x in w]
    return c, w
