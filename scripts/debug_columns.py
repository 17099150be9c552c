"""GPU debug helper: density / smoothed arrays of the column kernels vs the oracle on the F1 pages."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from multimodal_embeddings_b200 import ops, reference_api as api  # noqa: E402
from oracle import boxes as ob  # noqa: E402

pages = load_golden("f1_pages.json.gz")
f4 = load_golden("f4_stage45.json")
boxes = np.concatenate([np.asarray(p["boxes"], np.float64) for p in pages])
scores = np.concatenate([np.asarray(p["scores"], np.float64) for p in pages])
names = sum([p["class_names"] for p in pages], [])
flags = api._flags_from_names(names)
off = np.cumsum([0] + [len(p["boxes"]) for p in pages])
wh = [[p["image_size"]["width"], p["image_size"]["height"]] for p in pages]
med = np.asarray([g["median_width"] for g in f4])
centers, widths, n_cols, ws = ops.column_peaks(boxes, flags, scores, off, wh, med, 0.3, return_ws=True)
ws = ws.cpu().numpy()
centers, n_cols = centers.cpu().numpy(), n_cols.cpu().numpy()
for i, (p, g) in enumerate(zip(pages, f4)):
    w, h = wh[i]
    dens, res = ob.density_map(p["boxes"], p["class_names"], p["scores"], w, g["median_width"], 0.3)
    gw = ob.gaussian_window(g["median_width"], res)
    sm = np.convolve(dens, gw, mode="same")
    n = len(dens)
    dd = ws[i, 0, :n]
    ds = ws[i, 1, :n]
    bad = np.nonzero(dd != dens)[0]
    ok = [float(x) for x in centers[i, : n_cols[i]]] == g["column_centers"]
    print(i, p["name"][:20], "W", w, "nbins", n, "density mismatches", len(bad), bad[:8],
          "max|dsm|", float(np.abs(ds - sm).max()), "centers ok", ok)
    if len(bad):
        j = bad[0]
        print("   first bad bin", j, "gpu", dd[j], "ref", dens[j], "diff", dd[j] - dens[j])
