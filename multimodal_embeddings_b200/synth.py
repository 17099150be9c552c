"""Deterministic synthetic inputs for the page-geometry path (numpy, host side).

The reference ships no stage-1/2 outputs and its page scans are missing blobs
(SURVEY.md §4), so detections are synthesised: newspaper-like columns of stacked
regions, re-detected once per overlapping grid cell with jitter — the situation
2_edge_box_filter.py and 3_combine_grids.py exist to clean up.  Everything is a
pure function of (shape, grid, seed); ``seed = PAGE_SEED0 + page_idx`` makes the
corpus independent of how pages are sharded over GPUs.
"""
from __future__ import annotations

import numpy as np

PAGE_SEED0 = 0xB200

# DocLayout-YOLO id -> name (1_doclayout_bboxes.py:64-75)
ID_TO_NAMES = {
    0: "title", 1: "plain_text", 2: "abandon", 3: "figure", 4: "figure_caption",
    5: "table", 6: "table_caption", 7: "table_footnote", 8: "isolate_formula", 9: "formula_caption",
}
# empirical class mix of the reference's committed stage-3 outputs
_CLASS_IDS = np.array([1, 0, 3, 2, 4, 8, 5, 6, 9], dtype=np.int64)
_CLASS_P = np.array([2766, 822, 200, 153, 44, 28, 26, 6, 5], dtype=np.float64)
_CLASS_P /= _CLASS_P.sum()

# the 19 page sizes (W, H) of the reference's newspaper_images, in sorted-name order
FIXTURE_PAGE_SIZES = [
    (7934, 5755), (3631, 5285), (4029, 5180), (2928, 3951), (3861, 5264), (3946, 6000), (3960, 6010),
    (3784, 5045), (3837, 5153), (3940, 5824), (4039, 5745), (3798, 5120), (3801, 5601), (3989, 5222),
    (3500, 5266), (4170, 5407), (4000, 5443), (2778, 4187), (3400, 4864),
]


def grid_cells_f64(width, height, rows, cols, overlap_percentage):
    """Cell coordinates [rows*cols, 4] (x_start,y_start,x_end,y_end), float64, row-major.
    Same arithmetic as 1_doclayout_bboxes.py:388-421 (python floats are doubles)."""
    bw = width / cols
    bh = height / rows
    ox = bw * (overlap_percentage / 100)
    oy = bh * (overlap_percentage / 100)
    out = np.empty((rows * cols, 4), np.float64)
    for r in range(rows):
        for c in range(cols):
            xs = c * bw - (ox if c > 0 else 0)
            ys = r * bh - (oy if r > 0 else 0)
            xe = (c + 1) * bw + (ox if c < cols - 1 else 0)
            ye = (r + 1) * bh + (oy if r < rows - 1 else 0)
            out[r * cols + c] = (max(0, xs), max(0, ys), min(width, xe), min(height, ye))
    return out


def true_regions(width, height, rng, n_columns=None, density=1.0):
    """Column-structured 'true' layout: n_columns text columns of stacked regions."""
    if n_columns is None:
        n_columns = max(2, int(round(width / 1000)))
    margin = 0.02 * width
    gutter = 0.012 * width
    col_w = (width - 2 * margin - (n_columns - 1) * gutter) / n_columns
    boxes, classes = [], []
    for c in range(n_columns):
        x0 = margin + c * (col_w + gutter)
        y = 0.02 * height + rng.uniform(0, 40)
        while y < 0.98 * height:
            h = float(rng.choice([rng.uniform(18, 60), rng.uniform(60, 260), rng.uniform(260, 700)],
                                 p=[0.35, 0.5, 0.15])) / max(density, 1e-6)
            h = max(h, 6.0)
            cls = int(rng.choice(_CLASS_IDS, p=_CLASS_P))
            span = 1
            if cls in (0, 3, 5) and c < n_columns - 1 and rng.random() < 0.08:
                span = 2  # headline / figure spanning two columns
            w = span * col_w + (span - 1) * gutter
            inset = rng.uniform(0, 0.03) * col_w
            boxes.append((x0 + inset, y, x0 + w - inset, min(y + h, height - 1.0)))
            classes.append(cls)
            y += h + rng.uniform(2, 14) / max(density, 1e-6)
    return np.asarray(boxes, np.float64), np.asarray(classes, np.int64)


def page_detections(width, height, rows, cols, overlap_percentage, n_boxes, seed,
                    dups=2, jitter=2.0):
    """Per-tile detections for one page, in the layout stage 1 writes
    (1_doclayout_bboxes.py:576-589): cell-local float32 boxes + the cell they came
    from.  Returns dict with
      cells       [C,4] f64     page-coordinate cell rectangles
      box_cell    [N]  i32      cell index of each detection (non-decreasing)
      boxes_local [N,4] f64     float32-valued detector output, cell-local
      scores      [N]  f64      float32-valued
      classes     [N]  f64      float class ids (JSON floats in the reference)
    Exactly ``n_boxes`` detections (the layout density is scaled to reach it)."""
    rng = np.random.default_rng(seed)
    cells = grid_cells_f64(width, height, rows, cols, overlap_percentage)
    cover = float(((cells[:, 2] - cells[:, 0]) * (cells[:, 3] - cells[:, 1])).sum() / (width * height))
    # regions per page at density 1 ~= n_columns * H / 160
    ncol = max(2, int(round(width / 1000)))
    base = ncol * height / 160.0
    density = max(0.25, n_boxes * 1.25 / (base * cover * dups))
    tb, tc = true_regions(width, height, rng, ncol, density)
    loc, cell_of, cls = [], [], []
    for ci, (cx0, cy0, cx1, cy1) in enumerate(cells):
        ix0 = np.maximum(tb[:, 0], cx0)
        iy0 = np.maximum(tb[:, 1], cy0)
        ix1 = np.minimum(tb[:, 2], cx1)
        iy1 = np.minimum(tb[:, 3], cy1)
        vis = np.clip(ix1 - ix0, 0, None) * np.clip(iy1 - iy0, 0, None)
        area = (tb[:, 2] - tb[:, 0]) * (tb[:, 3] - tb[:, 1])
        sel = np.nonzero(vis > 0.3 * area)[0]
        for _ in range(dups):
            m = sel[rng.random(len(sel)) < 0.9]
            b = np.stack([ix0[m] - cx0, iy0[m] - cy0, ix1[m] - cx0, iy1[m] - cy0], 1)
            b = b + rng.normal(0, jitter, b.shape)
            b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, cx1 - cx0)
            b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, cy1 - cy0)
            b[:, 2] = np.maximum(b[:, 2], b[:, 0] + 1.0)
            b[:, 3] = np.maximum(b[:, 3], b[:, 1] + 1.0)
            c = tc[m].copy()
            flip = rng.random(len(m)) < 0.03
            c[flip] = rng.choice(_CLASS_IDS, size=int(flip.sum()), p=_CLASS_P)
            loc.append(b)
            cell_of.append(np.full(len(m), ci, np.int32))
            cls.append(c)
    loc = np.concatenate(loc) if loc else np.zeros((0, 4))
    cell_of = np.concatenate(cell_of) if cell_of else np.zeros((0,), np.int32)
    cls = np.concatenate(cls) if cls else np.zeros((0,), np.int64)
    n = len(loc)
    if n > n_boxes:
        keep = np.sort(rng.choice(n, n_boxes, replace=False))
    elif n < n_boxes:  # top up with re-jittered copies (keeps cell ordering)
        extra = rng.choice(max(n, 1), n_boxes - n, replace=True) if n else np.zeros(0, np.int64)
        keep = np.sort(np.concatenate([np.arange(n), extra]))
    else:
        keep = np.arange(n)
    loc, cell_of, cls = loc[keep], cell_of[keep], cls[keep]
    loc = loc + rng.normal(0, 0.25, loc.shape) * (np.arange(len(loc))[:, None] >= 0)
    loc = np.maximum(loc, 0.0).astype(np.float32).astype(np.float64)
    scores = rng.uniform(0.1, 0.99, len(loc)).astype(np.float32).astype(np.float64)
    return {
        "width": int(width), "height": int(height),
        "cells": cells,
        "box_cell": cell_of.astype(np.int32),
        "boxes_local": np.ascontiguousarray(loc),
        "scores": scores,
        "classes": cls.astype(np.float64),
    }


def class_names_of(classes):
    return [ID_TO_NAMES[int(c)] for c in classes]


def page_pixels(width, height, seed):
    """Newspaper-like uint8 BGR page (host twin of the device generator is not
    needed: tests upload this array).  Light background, dark 'ink' runs."""
    rng = np.random.default_rng(seed ^ 0x5EED)
    bg = rng.normal(225, 6, (height, width, 1))
    ink = rng.random((height, width, 1)) < 0.15
    img = np.where(ink, rng.normal(30, 10, (height, width, 1)), bg)
    img = img + rng.normal(0, 3, (height, width, 3))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def newspaper_page(width, height, seed):
    """Grey uint8 [H, W] scan-like page for the compressed-input leg (bench e2e, decoder tests): text lines of
    glyph-sized ink cells in columns with gutters, word gaps and paragraph breaks, rendered at a quarter of the
    resolution and enlarged bilinearly (soft stroke edges), on paper with slow shading and optically blurred
    grain.  Unlike `page_pixels` (independent noise per pixel, incompressible) it compresses like a scan: about
    3 bits/pixel as JPEG at cv2's default quality 95, 1.4 at quality 75."""
    import cv2
    rng = np.random.default_rng(seed)
    cw, ch = (width + 3) // 4, (height + 3) // 4
    ink = rng.random((ch, cw)) < 0.42
    line = (np.arange(ch) % 10 >= 1) & (np.arange(ch) % 10 <= 6)
    para = (np.arange(ch) // 10) % 9 == 8
    ncol = max(2, round(width / 1000))
    gutter = (np.arange(cw) % (cw / ncol)) < (cw / ncol) * 0.04
    gaps = np.zeros(cw, bool)
    pos = 0
    while pos < cw:
        pos += int(rng.integers(5, 14))
        gaps[pos:pos + 2] = True
        pos += 2
    mask = ink & (line & ~para)[:, None] & ~(gutter | gaps)[None, :]
    big = cv2.resize(np.where(mask, 40, 228).astype(np.uint8), (cw * 4, ch * 4), interpolation=cv2.INTER_LINEAR)
    big = big[:height, :width].astype(np.float32)
    gy, gx = np.mgrid[0:height + 64:64, 0:width + 64:64]
    shade = (6 * np.sin(gx / 900.0 + seed % 7) + 5 * np.cos(gy / 700.0)).astype(np.float32)
    big += cv2.resize(shade, (width, height), interpolation=cv2.INTER_LINEAR)
    big += cv2.blur(rng.integers(-6, 7, (height, width)).astype(np.float32), (3, 3))
    return np.clip(big, 0, 255).astype(np.uint8)
