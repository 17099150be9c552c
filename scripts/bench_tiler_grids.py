"""Tiler alone on the reference's DEFAULT grid set (1_doclayout_bboxes.py:718: full page + 2x2 + 3x3 + 4x4 =
30 tiles per page; SURVEY 8d third roofline row) and on other grid sets, 8000x6000 pages.  One JSON line per set.

    python scripts/bench_tiler_grids.py [--pages 32] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_embeddings_b200 import ops, synth  # noqa: E402

PEAK = 6533.8
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    with open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) as f:
        PEAK = json.load(f).get("hbm_gbs", PEAK)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="", help="substring of the grid-set name to run alone (ncu captures)")
    a = ap.parse_args()
    w, h = 8000, 6000
    sets = {"full+2x2+3x3+4x4 (reference default)": [(1, 1), (2, 2), (3, 3), (4, 4)], "full page only": [(1, 1)],
            "2x2": [(2, 2)], "3x3": [(3, 3)], "4x4 (cfg3)": [(4, 4)]}
    pages = None
    for name, grids in sets.items():
        if a.only and a.only not in name:
            continue
        plan = ops.TilePlan(w, h, grids, 20.0)
        if pages is None:
            pages = plan.alloc_pages(a.pages)
            ops.synth_pages(plan, a.pages, synth.PAGE_SEED0, out=pages)
        out = plan.alloc_out(a.pages)
        for _ in range(a.warmup):
            plan.run(pages, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            plan.run(pages, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        alg = plan.algorithmic_bytes * a.pages
        print(json.dumps({"grids": name, "tiles_per_page": len(plan.tiles), "pages": a.pages, "ms_per_launch": ms,
                          "pages_per_s": a.pages / ms * 1e3, "algorithmic_gb_per_launch": alg / 1e9,
                          "achieved_gbs": alg / ms / 1e6, "frac_of_measured_hbm_peak": alg / ms / 1e6 / PEAK}))
        del out, plan


if __name__ == "__main__":
    main()
