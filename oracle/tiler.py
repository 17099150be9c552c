"""Oracle for the tiler: overlapped grid geometry + letterbox resize/normalise.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call sites:
  * grid geometry ............ 1_doclayout_bboxes.py:366-444 (split_image_into_grid)
  * coordinate translation ... 1_doclayout_bboxes.py:484-511
  * grid config parsing ...... 1_doclayout_bboxes.py:656-680
  * detector front half ...... 1_doclayout_bboxes.py:191-210 -> third-party
    ``YOLOv10.predict(image, imgsz=1024)`` (doclayout-yolo / ultralytics, both
    unpinned and absent offline).  Its published pre-processing is
    LetterBox(new_shape=imgsz, auto, stride=32, center, scaleup) ->
    cv2.resize(INTER_LINEAR) -> cv2.copyMakeBorder(114) -> BGR->RGB -> HWC->CHW
    -> /255.  We restate that arithmetic here; pixels are pinned against the
    in-container cv2 4.13 primitives.
"""
from __future__ import annotations

import numpy as np

PAD_VALUE = 114
COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS  # cv2 INTER_RESIZE_COEF_SCALE


# ----------------------------------------------------------------------------
# grid geometry (1_doclayout_bboxes.py:366-444)
# ----------------------------------------------------------------------------
def grid_cells(width: int, height: int, rows: int, cols: int, overlap_percentage: float):
    """Cell coordinates for an overlapped rows x cols grid, row-major, 1-indexed.

    Follows 1_doclayout_bboxes.py:388-442: base cell = W/cols x H/rows (float),
    overlap = base * pct/100 added on *internal* sides only, clamp with
    ``max(0, .)`` / ``min(W, .)`` (which returns the *int* bound when it wins,
    so the JSON mixes ints and floats), slice with ``int()`` truncation.
    Returns a list of dicts: coordinates (un-truncated), slice (ints), row, col.
    """
    bw = width / cols
    bh = height / rows
    ox = bw * (overlap_percentage / 100)
    oy = bh * (overlap_percentage / 100)
    out = []
    for r in range(rows):
        for c in range(cols):
            xs = c * bw - (ox if c > 0 else 0)
            ys = r * bh - (oy if r > 0 else 0)
            xe = (c + 1) * bw + (ox if c < cols - 1 else 0)
            ye = (r + 1) * bh + (oy if r < rows - 1 else 0)
            xs = max(0, xs)
            ys = max(0, ys)
            xe = min(width, xe)
            ye = min(height, ye)
            out.append({
                "coordinates": {"x_start": xs, "y_start": ys, "x_end": xe, "y_end": ye},
                "slice": (int(xs), int(ys), int(xe), int(ye)),
                "row": r + 1,
                "col": c + 1,
            })
    return out


def split_array_into_grid(image: np.ndarray, rows: int, cols: int, overlap_percentage: float):
    """In-memory twin of split_image_into_grid (1_doclayout_bboxes.py:366-444) for
    an already-decoded BGR page (the reference decodes with cv2.imread at :381)."""
    h, w = image.shape[:2]
    cells = grid_cells(w, h, rows, cols, overlap_percentage)
    for cell in cells:
        x0, y0, x1, y1 = cell["slice"]
        cell["image"] = image[y0:y1, x0:x1]
    return cells


def translate_boxes(boxes, cell_coordinates):
    """1_doclayout_bboxes.py:484-511: add the *float* cell origin in Python double."""
    xo = cell_coordinates["x_start"]
    yo = cell_coordinates["y_start"]
    return [[b[0] + xo, b[1] + yo, b[2] + xo, b[3] + yo] for b in boxes]


def parse_grid_configs(grid_str):
    """1_doclayout_bboxes.py:656-680: "2x2,3x3" -> [(2,2),(3,3)]; entries without
    'x' are skipped; a ValueError keeps what was parsed so far."""
    out = []
    try:
        if grid_str:
            for part in grid_str.split(","):
                part = part.strip()
                if "x" in part:
                    a, b = part.split("x")
                    out.append((int(a), int(b)))
    except ValueError:
        pass
    return out


# ----------------------------------------------------------------------------
# letterbox geometry (ultralytics LetterBox, published behaviour)
# ----------------------------------------------------------------------------
def letterbox_geometry(src_w: int, src_h: int, imgsz: int = 1024, stride: int = 32,
                       auto: bool = True, scaleup: bool = True):
    """Returns dict(new_w,new_h,pad_l,pad_t,out_w,out_h).

    r = min(S/h, S/w); new = (round(w r), round(h r)) with Python (half-even)
    rounding; dw,dh = S-new; auto -> mod stride; halve; top=round(dh-0.1),
    bottom=round(dh+0.1), same for left/right."""
    r = min(imgsz / src_h, imgsz / src_w)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = int(round(src_w * r)), int(round(src_h * r))
    dw, dh = imgsz - new_w, imgsz - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return {"new_w": new_w, "new_h": new_h, "pad_l": left, "pad_t": top,
            "out_w": new_w + left + right, "out_h": new_h + top + bottom}


# ----------------------------------------------------------------------------
# pixels
# ----------------------------------------------------------------------------
def letterbox_tile_cv2(cell_bgr: np.ndarray, imgsz: int = 1024, stride: int = 32,
                       auto: bool = True, scaleup: bool = True) -> np.ndarray:
    """The reference's own primitives on one tile: cv2.resize(INTER_LINEAR) +
    cv2.copyMakeBorder(114) + BGR->RGB + CHW.  Returns uint8 [3, out_h, out_w]."""
    import cv2
    h, w = cell_bgr.shape[:2]
    g = letterbox_geometry(w, h, imgsz, stride, auto, scaleup)
    img = cell_bgr
    if (w, h) != (g["new_w"], g["new_h"]):
        img = cv2.resize(img, (g["new_w"], g["new_h"]), interpolation=cv2.INTER_LINEAR)
    bottom = g["out_h"] - g["new_h"] - g["pad_t"]
    right = g["out_w"] - g["new_w"] - g["pad_l"]
    img = cv2.copyMakeBorder(img, g["pad_t"], bottom, g["pad_l"], right,
                             cv2.BORDER_CONSTANT, value=(PAD_VALUE,) * 3)
    return np.ascontiguousarray(img[..., ::-1].transpose(2, 0, 1))


def u8_to_f16_unit(x_u8: np.ndarray) -> np.ndarray:
    """uint8 -> fp16 in [0,1]: fp16(float32(v)/255) (ultralytics ``im.half()/255``
    evaluates the division in fp32 and rounds once to fp16)."""
    return (x_u8.astype(np.float32) / np.float32(255.0)).astype(np.float16)


def resize_coeffs(ssize: int, dsize: int, clamp_frac: bool):
    """cv2 INTER_LINEAR uint8 coefficient tables (validated 100 % exact
    against cv2 4.13 on random up/down-scales):
      scale = 1/(dsize/ssize) in double; f = float32((d+0.5)*scale-0.5);
      s = floor(f); f -= s (float32).  x-direction clamps (s<0 -> s=0,f=0;
      s>=ssize-1 -> s=ssize-1,f=0); y-direction keeps f and clamps the two row
      indices separately.  c0 = rint((1-f)*2048), c1 = rint(f*2048) in float32."""
    inv = float(dsize) / float(ssize)
    scale = 1.0 / inv
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_frac:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= ssize - 1
        f[hi] = 0
        s[hi] = ssize - 1
    c0 = np.rint((np.float32(1.0) - f).astype(np.float32) * np.float32(COEF_ONE)).astype(np.int32)
    c1 = np.rint(f * np.float32(COEF_ONE)).astype(np.int32)
    s0 = np.clip(s, 0, ssize - 1).astype(np.int32)
    s1 = np.clip(s + 1, 0, ssize - 1).astype(np.int32)
    return s0, s1, c0, c1


def resize_fixed_point(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """numpy restatement of cv2.resize(uint8, INTER_LINEAR): 11-bit horizontal pass
    h = S[x0]*a0 + S[x1]*a1, vertical (((b0*(h0>>4))>>16)+((b1*(h1>>4))>>16)+2)>>2."""
    sh, sw = img.shape[:2]
    x0, x1, a0, a1 = resize_coeffs(sw, dw, True)
    y0, y1, b0, b1 = resize_coeffs(sh, dh, False)
    src = img.astype(np.int32)
    h = src[:, x0, :] * a0[None, :, None] + src[:, x1, :] * a1[None, :, None]
    h0 = h[y0] >> 4
    h1 = h[y1] >> 4
    out = (((b0[:, None, None] * h0) >> 16) + ((b1[:, None, None] * h1) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_tile_model(cell_bgr: np.ndarray, imgsz: int = 1024, stride: int = 32,
                         auto: bool = True, scaleup: bool = True) -> np.ndarray:
    """Same as letterbox_tile_cv2 but with the numpy fixed-point resize (no cv2)."""
    h, w = cell_bgr.shape[:2]
    g = letterbox_geometry(w, h, imgsz, stride, auto, scaleup)
    rs = resize_fixed_point(cell_bgr, g["new_w"], g["new_h"])
    out = np.full((g["out_h"], g["out_w"], 3), PAD_VALUE, np.uint8)
    out[g["pad_t"]:g["pad_t"] + g["new_h"], g["pad_l"]:g["pad_l"] + g["new_w"]] = rs
    return np.ascontiguousarray(out[..., ::-1].transpose(2, 0, 1))


def tile_page(page_bgr: np.ndarray, rows: int, cols: int, overlap_percentage: float,
              imgsz: int = 1024, stride: int = 32, auto: bool = True, use_cv2: bool = True):
    """Whole front half for one page: list of (cell dict, fp16 CHW tile)."""
    fn = letterbox_tile_cv2 if use_cv2 else letterbox_tile_model
    out = []
    for cell in split_array_into_grid(page_bgr, rows, cols, overlap_percentage):
        out.append((cell, u8_to_f16_unit(fn(cell["image"], imgsz, stride, auto))))
    return out
