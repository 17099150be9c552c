"""CPU stand-ins for the device wrappers of multimodal_embeddings_b200.ops, built on the oracle.

TEST INFRASTRUCTURE.  The command lines (cli.py) are host logic around kernel launches; to check that logic —
file discovery, pooling order, schemas, key order, error handling — against the reference's golden trees on a
machine WITHOUT a GPU, `install(monkeypatch)` swaps the wrappers the command lines call for functions with the
same signatures that evaluate the oracle (oracle/boxes.py, oracle/tiler.py).  Nothing here ships: the product has
no CPU path, and the `-m gpu` twins of these tests run the real kernels on the same goldens.
"""
from __future__ import annotations

import numpy as np
import torch

from multimodal_embeddings_b200 import _lib, ops
from oracle import boxes as ob
from oracle.nms_fast import nms_pick_order_c


def _np(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype))


def edge_filter(boxes, box_cell, cells, page_wh, page_off, threshold=10, boxes_are_local=True,
                want_boxes_page=True, stream=None):
    boxes = _np(boxes, np.float64).reshape(-1, 4)
    box_cell, cells = _np(box_cell, np.int32), _np(cells, np.float64).reshape(-1, 4)
    page_wh, page_off = _np(page_wh, np.int32).reshape(-1, 2), _np(page_off, np.int64)
    n, p = len(boxes), len(page_wh)
    bp = boxes.copy()
    if boxes_are_local:
        bp = boxes + cells[box_cell][:, [0, 1, 0, 1]]
    keep = np.zeros(n, np.uint8)
    kept_idx = np.zeros(max(n, 1), np.int32)
    n_kept = np.zeros(max(p, 1), np.int32)
    for pg in range(p):
        w, h = int(page_wh[pg, 0]), int(page_wh[pg, 1])
        k = 0
        for i in range(int(page_off[pg]), int(page_off[pg + 1])):
            if not ob.touches_internal_edge(bp[i].tolist(), cells[box_cell[i]].tolist(), w, h, threshold):
                keep[i] = 1
                kept_idx[page_off[pg] + k] = i
                k += 1
        n_kept[pg] = k
    t = torch.from_numpy
    return (t(bp) if want_boxes_page else None), t(keep), t(kept_idx), t(n_kept[:p])


class _Ws:
    def stats(self):
        return {"status": 0, "candidate_block_pairs": 0, "rounds": 0, "box_pairs_tested": 0}


def NmsWorkspace(*a, **k):
    return _Ws()


def nms_merge(boxes, scores, classes, page_off, iou_threshold=0.5, sel_idx=None, n_sel=None, max_boxes_per_page=0,
              workspace=None, stream=None, kept_idx=None, n_kept=None, mode=0, **_):
    assert sel_idx is None
    boxes, scores = _np(boxes, np.float64).reshape(-1, 4), _np(scores, np.float64)
    page_off = _np(page_off, np.int64)
    n, p = len(boxes), len(page_off) - 1
    kept = np.zeros(max(n, 1), np.int32)
    nk = np.zeros(max(p, 1), np.int32)
    for pg in range(p):
        a, b = int(page_off[pg]), int(page_off[pg + 1])
        if mode & _lib.PG_NMS_FP32:
            order = ob.nms_torchvision_f32(boxes[a:b].astype(np.float32), scores[a:b].astype(np.float32), iou_threshold)
        else:
            order = nms_pick_order_c(boxes[a:b], scores[a:b], _np(classes, np.float64)[a:b], iou_threshold)
        order = np.asarray(order, np.int64)
        kept[a:a + len(order)] = order + a
        nk[pg] = len(order)
    return torch.from_numpy(kept), torch.from_numpy(nk[:p]), _Ws()


def width_median(boxes, flags, page_off, page_wh, min_margin_percent=0.2, sel_idx=None, n_sel=None, width_hist=None,
                 stream=None):
    boxes, flags = _np(boxes, np.float64).reshape(-1, 4), _np(flags, np.uint8)
    page_off, page_wh = _np(page_off, np.int64), _np(page_wh, np.int32).reshape(-1, 2)
    p = len(page_wh)
    med, nb = np.zeros(max(p, 1)), np.zeros(max(p, 1), np.int32)
    for pg in range(p):
        a, b = int(page_off[pg]), int(page_off[pg + 1])
        widths = [bb[2] - bb[0] for bb, f in zip(boxes[a:b].tolist(), flags[a:b]) if f & _lib.PG_FLAG_PLAIN_TEXT]
        bins = ob.bin_widths(widths, min_margin_percent, int(page_wh[pg, 0]))
        nb[pg] = len(bins)
        med[pg] = float(ob.median_of_bins(bins))
    return torch.from_numpy(med[:p]), torch.from_numpy(nb[:p])


def column_peaks(boxes, flags, scores, page_off, page_wh, median, min_confidence=0.3, sel_idx=None, n_sel=None,
                 max_cols=64, max_bins=0, col_hist=None, stream=None, return_ws=False):
    boxes, flags = _np(boxes, np.float64).reshape(-1, 4), _np(flags, np.uint8)
    scores, page_off = _np(scores, np.float64), _np(page_off, np.int64)
    page_wh, median = _np(page_wh, np.int32).reshape(-1, 2), _np(median, np.float64)
    p = len(page_wh)
    centers, widths = np.zeros((max(p, 1), max_cols), np.int32), np.zeros((max(p, 1), max_cols))
    n_cols = np.zeros(max(p, 1), np.int32)
    for pg in range(p):
        a, b = int(page_off[pg]), int(page_off[pg + 1])
        names = ["plain_text" if f & _lib.PG_FLAG_PLAIN_TEXT else "title" if f & _lib.PG_FLAG_TITLE else "x"
                 for f in flags[a:b]]
        c, w = ob.column_centers(boxes[a:b].tolist(), names, scores[a:b].tolist(), int(page_wh[pg, 0]),
                                 int(page_wh[pg, 1]), float(median[pg]), min_confidence)
        n_cols[pg] = len(c)
        centers[pg, :len(c)] = c
        widths[pg, :len(w)] = w
    return torch.from_numpy(centers[:p]), torch.from_numpy(widths[:p]), torch.from_numpy(n_cols[:p])


class TileBatch:
    """Plans only (pg_tile_plan_* is host code); no pixels are produced — the tests that use this stand-in
    plug in detectors that do not look at the tiles."""

    def __init__(self, sizes, grids=((2, 2),), overlap=20.0, imgsz=1024, stride=32, auto=True, channels=3):
        self.sizes = [(int(w), int(h)) for w, h in sizes]
        self.plans = {s: ops.TilePlan(s[0], s[1], grids, overlap, imgsz, stride, auto, channels=channels) for s in set(self.sizes)}

    def plan_of(self, page):
        return self.plans[self.sizes[page]]

    def bind(self, pages, stream=None):
        return []

    def run(self, stream=None):
        return []

    def tile_view(self, page, tile):
        return None


def install(monkeypatch):
    """Swap the device wrappers for the oracle stand-ins and route all JSON through CPython."""
    monkeypatch.setenv("PG_PYTHON_JSON", "1")
    for name in ("edge_filter", "nms_merge", "width_median", "column_peaks", "TileBatch", "NmsWorkspace"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, "upload_pages_pinned", lambda images, stream=None: list(images))
