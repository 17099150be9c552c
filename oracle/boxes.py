"""Oracle for the box stages: edge-touch filter, class-aware greedy NMS, width
median, column centres.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

All arithmetic is Python double / numpy float64, evaluated in the reference's
operation order, because the kernels are required to be bit-exact against it.
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------------------
# stage 2 — edge-touch filter (2_edge_box_filter.py)
# ----------------------------------------------------------------------------
def touches_internal_edge(box, cell, image_width, image_height, threshold=10):
    """2_edge_box_filter.py:44-90.  ``cell`` is (x_start,y_start,x_end,y_end) in
    page coordinates.  A side is *internal* when it is more than ``threshold`` px
    from the page border; a box is dropped when it reaches within ``threshold``
    of an internal side.  Test order right, bottom, left, top (:71-88)."""
    x_min, y_min, x_max, y_max = box
    cx0, cy0, cx1, cy1 = cell
    if abs(cx1 - image_width) > threshold and x_max >= (cx1 - threshold):
        return True
    if abs(cy1 - image_height) > threshold and y_max >= (cy1 - threshold):
        return True
    if cx0 > threshold and x_min <= (cx0 + threshold):
        return True
    if cy0 > threshold and y_min <= (cy0 + threshold):
        return True
    return False


def cell_tuple(cell_coordinates, image_width, image_height):
    """dict/list handling of 2_edge_box_filter.py:62-68 (dict defaults 0/W/H)."""
    if isinstance(cell_coordinates, dict):
        return (cell_coordinates.get("x_start", 0), cell_coordinates.get("y_start", 0),
                cell_coordinates.get("x_end", image_width), cell_coordinates.get("y_end", image_height))
    return tuple(cell_coordinates)


def filter_cells(cells, image_width, image_height, threshold=10):
    """Index form of filter_grid_info (2_edge_box_filter.py:206-217): for each cell
    (dict with 'cell_coordinates' and 'boxes_original') return the kept positions,
    in input order."""
    kept = []
    for cell in cells:
        ct = cell_tuple(cell["cell_coordinates"], image_width, image_height)
        kept.append([i for i, b in enumerate(cell["boxes_original"])
                     if not touches_internal_edge(b, ct, image_width, image_height, threshold)])
    return kept


# ----------------------------------------------------------------------------
# stage 3 — IoU + class-aware greedy NMS (3_combine_grids.py)
# ----------------------------------------------------------------------------
def iou(a, b):
    """3_combine_grids.py:46-78.  Touching/disjoint -> 0.0; union<=0 -> 0.0;
    union = area_a + area_b - inter evaluated left to right, one IEEE divide."""
    xl = max(a[0], b[0])
    yt = max(a[1], b[1])
    xr = min(a[2], b[2])
    yb = min(a[3], b[3])
    if xr < xl or yb < yt:
        return 0.0
    inter = (xr - xl) * (yb - yt)
    area_a = (a[2] - a[0]) * (a[3] - a[1])
    area_b = (b[2] - b[0]) * (b[3] - b[1])
    union = area_a + area_b - inter
    return inter / union if union > 0 else 0.0


def nms_pick_order(boxes, scores, classes, iou_threshold=0.5):
    """Index form of apply_non_max_suppression (3_combine_grids.py:80-138).

    Repeatedly pick the *first* remaining box with the maximum score (:112), emit
    it, then drop every remaining box with ``iou(pick, box) > thr`` and the same
    class (:130).  Returns the picked input positions in pick order.  Pure-Python
    O(N*K) like the reference (this is the timed CPU baseline for the merge)."""
    alive = list(range(len(boxes)))
    picks = []
    while alive:
        alive_scores = [scores[i] for i in alive]
        pos = alive_scores.index(max(alive_scores))
        cur = alive.pop(pos)
        picks.append(cur)
        cb, cc = boxes[cur], classes[cur]
        alive = [j for j in alive if not (iou(cb, boxes[j]) > iou_threshold and classes[j] == cc)]
    return picks


def pool_boxes(docs):
    """Concatenation order of combine_boxes_for_image (3_combine_grids.py:222-267):
    documents in the given order; a grid-info document contributes each cell's
    ``boxes_original`` in cell order, a standard/per-cell document contributes
    ``boxes_original`` if present else ``boxes``.  Returns boxes, scores, classes,
    class_names, image_path, image_size (image_size only from standard docs)."""
    boxes, scores, classes, names = [], [], [], []
    image_path = None
    image_size = None
    for d in docs:
        if "cells" in d:
            if not image_path and "original_image_path" in d:
                image_path = d["original_image_path"]
            for cell in d["cells"]:
                reg = cell.get("regions")
                if reg is not None and "boxes_original" in reg:
                    boxes.extend(reg["boxes_original"])
                    scores.extend(reg["scores"])
                    classes.extend(reg["classes"])
                    names.extend(reg["class_names"])
        elif "boxes" in d:
            if not image_path and "image_path" in d:
                image_path = d["image_path"]
            if not image_size and "image_size" in d:
                image_size = d["image_size"]
            boxes.extend(d["boxes_original"] if "boxes_original" in d else d["boxes"])
            scores.extend(d["scores"])
            classes.extend(d["classes"])
            names.extend(d["class_names"])
    return boxes, scores, classes, names, image_path, image_size


# ----------------------------------------------------------------------------
# stage 4 — width bins + median (4_extract_median_widths.py)
# ----------------------------------------------------------------------------
def bin_widths(widths, min_margin_percent, page_width):
    """4_extract_median_widths.py:49-80: sequential leader binning.  A width joins
    the *smallest-key* existing bin with |w-key| <= margin, else founds a bin keyed
    by itself (``bins[w] = 1`` — which *resets* an equal key when margin < 0)."""
    if not widths:
        return {}
    margin = page_width * (min_margin_percent / 100)
    bins = {}
    for w in widths:
        for key in sorted(bins):
            if abs(w - key) <= margin:
                bins[key] += 1
                break
        else:
            bins[w] = 1
    return bins


def median_of_bins(bins):
    """4_extract_median_widths.py:82-101: np.median over keys repeated by count;
    empty -> int 0."""
    if not bins:
        return 0
    flat = []
    for key, cnt in bins.items():
        flat.extend([key] * cnt)
    return np.median(flat)


def plain_text_widths(boxes, class_names):
    """4_extract_median_widths.py:135-141."""
    return [boxes[i][2] - boxes[i][0] for i, n in enumerate(class_names)
            if n == "plain_text" and i < len(boxes)]


def median_width(boxes, class_names, page_width, min_margin_percent=0.2):
    bins = bin_widths(plain_text_widths(boxes, class_names), min_margin_percent, page_width)
    return median_of_bins(bins), len(bins)


# ----------------------------------------------------------------------------
# stage 5 — column centres (5_detect_column_centers.py)
# ----------------------------------------------------------------------------
def local_maxima(x):
    """scipy.signal._peak_finding_utils._local_maxima_1d: strict rise,
    plateau midpoint (floor), first/last sample never a peak."""
    n = len(x)
    out = []
    i = 1
    while i < n - 1:
        if x[i - 1] < x[i]:
            a = i + 1
            while a < n - 1 and x[a] == x[i]:
                a += 1
            if x[a] < x[i]:
                out.append((i + a - 1) // 2)
                i = a
        i += 1
    return out


def select_by_distance(peaks, heights, distance):
    """scipy _select_by_peak_distance: highest first; drop neighbours closer than
    ceil(distance).  Priority ties resolve like a stable ascending argsort walked
    from the end (the later peak wins)."""
    d = math.ceil(distance)
    n = len(peaks)
    keep = [True] * n
    order = sorted(range(n), key=lambda k: heights[k])  # stable ascending
    for j in reversed(order):
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < d:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < n and peaks[k] - peaks[j] < d:
            keep[k] = False
            k += 1
    return keep


def prominence(x, peak):
    """scipy _peak_prominences with wlen=None."""
    left_min = x[peak]
    i = peak
    while i >= 0 and x[i] <= x[peak]:
        if x[i] < left_min:
            left_min = x[i]
        i -= 1
    right_min = x[peak]
    i = peak
    while i < len(x) and x[i] <= x[peak]:
        if x[i] < right_min:
            right_min = x[i]
        i += 1
    return x[peak] - max(left_min, right_min)


def find_peaks_restated(x, height, distance, prom):
    """find_peaks(x, height=, distance=, prominence=) in scipy's fixed step order:
    local maxima -> height -> distance -> prominence (5_detect_column_centers.py:164-169)."""
    peaks = [p for p in local_maxima(x) if height <= x[p]]
    keep = select_by_distance(peaks, [x[p] for p in peaks], distance)
    peaks = [p for p, k in zip(peaks, keep) if k]
    return [p for p in peaks if prom <= prominence(x, p)]


def gaussian_window(median_w, resolution):
    """5_detect_column_centers.py:147-153 — window length (odd, >=5), sigma=len/6,
    scipy.signal.windows.gaussian formula, normalised to sum 1."""
    m = max(5, int(median_w / (4 * resolution)))
    if m % 2 == 0:
        m += 1
    sigma = m / 6.0
    n = np.arange(0, m) - (m - 1.0) / 2.0
    w = np.exp(-(n ** 2) / (2 * sigma * sigma))
    return w / w.sum()


def density_map(boxes, class_names, scores, page_width, median_w, min_confidence=0.3):
    """5_detect_column_centers.py:109-144.  Returns (density, resolution) or
    (None, resolution) when no box passes the class/score filter."""
    resolution = max(1, int(page_width / 1000))
    num_bins = page_width // resolution + 1
    sel = [b for b, n, s in zip(boxes, class_names, scores)
           if n in ("plain_text", "title") and s >= min_confidence]
    if not sel:
        return None, resolution
    density = np.zeros(num_bins)
    for b in sel:
        x1, _, x2, _ = (int(v) for v in b)
        width = x2 - x1
        if 0.33 * median_w <= width <= 2.0 * median_w:
            left = max(0, x1 // resolution)
            right = min(num_bins - 1, x2 // resolution)
            center = (x1 + x2) // (2 * resolution)
            half = (right - left) / 2 + 1e-6
            for k in range(left, right + 1):
                density[k] += 1.0 - 0.5 * min(1.0, abs(k - center) / half)
    return density, resolution


def column_centers(boxes, class_names, scores, page_width, page_height, median_w,
                   min_confidence=0.3, use_scipy=True, return_debug=False):
    """find_column_centers (5_detect_column_centers.py:91-224): returns
    (centers, widths).  ``use_scipy`` picks scipy.signal.find_peaks (what the
    reference calls) or the restatement above."""
    density, res = density_map(boxes, class_names, scores, page_width, median_w, min_confidence)
    if density is None:
        return ([], []) if not return_debug else ([], [], None)
    g = gaussian_window(median_w, res)
    sm = np.convolve(density, g, mode="same")
    top = max(sm)
    hmin = top * 0.2
    dist = max(1, int(median_w / (1.5 * res)))
    if use_scipy:
        from scipy.signal import find_peaks
        peaks, _ = find_peaks(sm, height=hmin, distance=dist, prominence=top * 0.05)
        peaks = [int(p) for p in peaks]
    else:
        peaks = find_peaks_restated(sm, hmin, dist, top * 0.05)
    if len(peaks) == 0:
        return ([], []) if not return_debug else ([], [], sm)
    centers = [p * res for p in peaks]
    widths = valley_widths(sm, peaks, res, hmin, median_w)
    return (centers, widths) if not return_debug else (centers, widths, sm)


def valley_widths(sm, peaks, res, hmin, median_w):
    """5_detect_column_centers.py:179-222: per peak walk towards the neighbouring
    peaks tracking the strict minimum, stop after a sample below hmin*0.1; no
    minimum found -> midpoint; clamp w<0.5*med -> med, w>2.5*med -> 2*med."""
    out = []
    for i, p in enumerate(peaks):
        left = p
        if i > 0:
            prev = peaks[i - 1]
            for j in range(p - 1, prev, -1):
                if sm[j] < sm[left]:
                    left = j
                if sm[j] < hmin * 0.1:
                    break
            if left == p:
                left = (p + prev) // 2
        right = p
        if i < len(peaks) - 1:
            nxt = peaks[i + 1]
            for j in range(p + 1, nxt):
                if sm[j] < sm[right]:
                    right = j
                if sm[j] < hmin * 0.1:
                    break
            if right == p:
                right = (p + nxt) // 2
        w = (right - left) * res
        if w < 0.5 * median_w:
            w = median_w
        elif w > 2.5 * median_w:
            w = 2.0 * median_w
        out.append(w)
    return out


# ----------------------------------------------------------------------------
# stage 1 — per-tile class-agnostic NMS (torchvision.ops.nms at 1_doclayout_bboxes.py:217-225)
# ----------------------------------------------------------------------------
def nms_torchvision_f32(boxes, scores, iou_threshold):
    """Restatement of torchvision's CPU nms kernel (third-party, present in the image; pinned by goldens
    generated with torchvision 0.26): stable descending score order, everything in float32 —
    w = max(0, xx2-xx1), inter = w*h, ovr = inter / (area_i + area_j - inter) — suppress when
    ovr > threshold (threshold a double).  Returns kept indices in decreasing score order."""
    b = np.asarray(boxes, np.float32).reshape(-1, 4)
    s = np.asarray(scores, np.float32).reshape(-1)
    n = len(s)
    areas = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    order = np.argsort(-s.astype(np.float64), kind="stable")
    dead = np.zeros(n, bool)
    keep = []
    zero = np.float32(0)
    for a in range(n):
        i = order[a]
        if dead[i]:
            continue
        keep.append(int(i))
        rest = order[a + 1:]
        rest = rest[~dead[rest]]
        if len(rest) == 0:
            continue
        xx1 = np.maximum(b[i, 0], b[rest, 0])
        yy1 = np.maximum(b[i, 1], b[rest, 1])
        xx2 = np.minimum(b[i, 2], b[rest, 2])
        yy2 = np.minimum(b[i, 3], b[rest, 3])
        w = np.maximum(zero, xx2 - xx1)
        h = np.maximum(zero, yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        dead[rest[ovr.astype(np.float64) > iou_threshold]] = True
    return keep


# ----------------------------------------------------------------------------
# derived: per-box column assignment (no reference analogue, SURVEY.md §8a note)
# ----------------------------------------------------------------------------
def assign_columns(boxes, centers):
    """Index of the column centre nearest to each box's x-centre (x0+x1)/2, first minimum on ties;
    -1 when there are no centres.  Specification of pg_assign_columns."""
    out = []
    for b in boxes:
        cx = (b[0] + b[2]) / 2.0
        best, bd = -1, 0.0
        for c, cc in enumerate(centers):
            dd = abs(cx - float(cc))
            if best < 0 or dd < bd:
                best, bd = c, dd
        out.append(best)
    return out
