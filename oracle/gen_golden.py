"""Generate tests/golden/* by running the UNMODIFIED reference scripts.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs
/root/reference, which does not exist on the GPU box); the fixtures it writes
are committed and travel.

    python -m oracle.gen_golden            # regenerates every fixture

Fixtures
  f1_pages.json.gz     the reference's committed stage-3 outputs
                       (3_combined_bboxes/json/*.json), compacted: name, image_size,
                       boxes, classes, scores, class_names.  Data, not source.
  f4_stage45.json      reference bin_widths / calculate_median_width /
                       find_column_centers on every F1 page.
  stage1_geometry.json reference split_image_into_grid / translate_coordinates_to_original /
                       parse_grid_configs on blank pages of the fixture sizes.
  stage2_filter.json.gz reference is_box_touching_internal_edge + filter_grid_info on
                       synthetic per-tile detections (+ hand-made threshold-edge cases).
  stage3_nms.npz       reference apply_non_max_suppression on synthetic pooled boxes
                       (inputs + kept positions in pick order), incl. ties / degenerate boxes.
  stage45_synth.json   reference stage 4+5 on synthetic kept sets.
"""
from __future__ import annotations

import glob
import gzip
import hashlib
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from multimodal_embeddings_b200 import synth  # noqa: E402


def load_ref(stem):
    spec = importlib.util.spec_from_file_location("ref_" + stem.split("_")[0], os.path.join(REF, stem + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def f1_pages():
    pages = []
    for path in sorted(glob.glob(os.path.join(REF, "3_combined_bboxes", "json", "*.json"))):
        with open(path) as f:
            d = json.load(f)
        pages.append({
            "name": os.path.basename(path)[: -len("_combined.json")],
            "image_size": d["image_size"],
            "iou_threshold": d["parameters"]["iou_threshold"],
            "boxes": d["boxes"], "classes": d["classes"], "scores": d["scores"],
            "class_names": d["class_names"],
            "source_jsons": [os.path.basename(p) for p in d["source_jsons"]],
        })
    return pages


def gen_f4(pages, r4, r5):
    out = []
    for p in pages:
        w, h = p["image_size"]["width"], p["image_size"]["height"]
        widths = [b[2] - b[0] for b, n in zip(p["boxes"], p["class_names"]) if n == "plain_text"]
        bins = r4.bin_widths(widths, 0.2, w)
        med = r4.calculate_median_width(bins)
        centers, cw = r5.find_column_centers(p["boxes"], p["class_names"], p["scores"], w, h, med, 0.3)
        out.append({"name": p["name"], "n_plain_text": len(widths), "n_bins": len(bins),
                    "median_width": float(med), "column_centers": [float(c) for c in centers],
                    "column_widths": [float(x) for x in cw]})
    return out


def gen_stage1(r1):
    import cv2
    cases = []
    shapes = [(8000, 6000), (7934, 5755), (3801, 5601), (2778, 4187), (2928, 3951), (1001, 777), (640, 480)]
    grids = [(1, 1, 20.0), (2, 2, 20.0), (3, 3, 20.0), (4, 4, 20.0), (2, 3, 12.5), (4, 4, 0.0), (5, 2, 33.3)]
    with tempfile.TemporaryDirectory() as td:
        for (w, h) in shapes:
            path = os.path.join(td, f"p_{w}x{h}.png")
            cv2.imwrite(path, np.zeros((h, w), np.uint8))
            for (rows, cols, ov) in grids:
                cells = r1.split_image_into_grid(path, rows, cols, ov)
                cases.append({
                    "width": w, "height": h, "rows": rows, "cols": cols, "overlap": ov,
                    "cells": [{"coordinates": c["coordinates"], "row": c["row"], "col": c["col"],
                               "shape": list(c["image"].shape[:2])} for c in cells],
                })
    rng = np.random.default_rng(7)
    tr = []
    for _ in range(8):
        boxes = rng.uniform(0, 3000, (5, 4)).astype(np.float32).astype(np.float64).tolist()
        cc = {"x_start": float(rng.uniform(0, 5000)), "y_start": float(rng.uniform(0, 5000)),
              "x_end": 0.0, "y_end": 0.0}
        tr.append({"boxes": boxes, "cell_coordinates": cc,
                   "out": r1.translate_coordinates_to_original(boxes, cc)})
    pg = [{"in": s, "out": [list(t) for t in r1.parse_grid_configs(s)]}
          for s in ["2x2,3x3,4x4", "2x2", " 3x4 , 5x1", "", "7", "2x2,axb,3x3", "4x4,"]]
    return {"split": cases, "translate": tr, "parse_grid_configs": pg}


def gen_stage2(r2):
    import cv2
    cases = []
    cfgs = [(8000, 6000, 4, 4, 20.0, 3000, 10), (3801, 5601, 2, 2, 20.0, 2000, 10),
            (2778, 4187, 3, 3, 20.0, 1500, 10), (7934, 5755, 4, 4, 20.0, 2500, 25),
            (4000, 5443, 2, 2, 20.0, 800, 0), (640, 480, 2, 2, 20.0, 200, 10)]
    with tempfile.TemporaryDirectory() as td:
        for k, (w, h, rows, cols, ov, n, thr) in enumerate(cfgs):
            det = synth.page_detections(w, h, rows, cols, ov, n, synth.PAGE_SEED0 + 100 + k)
            cells = det["cells"]
            boxes_page = det["boxes_local"] + cells[det["box_cell"]][:, [0, 1, 0, 1]]
            # put some boxes exactly on the decision boundaries cell_edge -/+ thr
            rng = np.random.default_rng(k)
            for j in rng.choice(len(boxes_page), min(40, len(boxes_page)), replace=False):
                c = cells[det["box_cell"][j]]
                which = j % 4
                if which == 0:
                    boxes_page[j, 2] = c[2] - thr
                elif which == 1:
                    boxes_page[j, 3] = c[3] - thr
                elif which == 2:
                    boxes_page[j, 0] = c[0] + thr
                else:
                    boxes_page[j, 1] = c[1] + thr
            img = os.path.join(td, f"page{k}.png")
            cv2.imwrite(img, np.zeros((h, w), np.uint8))
            grid_info = {"original_image_path": img,
                         "grid_config": {"rows": rows, "cols": cols, "overlap_percentage": ov}, "cells": []}
            ref_cells = r1_cells(w, h, rows, cols, ov)
            for ci in range(rows * cols):
                m = np.nonzero(det["box_cell"] == ci)[0]
                grid_info["cells"].append({
                    "cell_path": f"c{ci}.png", "cell_json_path": f"c{ci}.json",
                    "cell_coordinates": ref_cells[ci], "row": ci // cols + 1, "col": ci % cols + 1,
                    "regions": {"boxes": det["boxes_local"][m].tolist(), "boxes_original": boxes_page[m].tolist(),
                                "classes": det["classes"][m].tolist(), "scores": det["scores"][m].tolist(),
                                "class_names": synth.class_names_of(det["classes"][m])}})
            filt = r2.filter_grid_info(grid_info, thr)
            kept = []
            for ci, cell in enumerate(grid_info["cells"]):
                pred = [bool(r2.is_box_touching_internal_edge(b, cell["cell_coordinates"], w, h, thr))
                        for b in cell["regions"]["boxes_original"]]
                kidx = [i for i, p in enumerate(pred) if not p]
                assert filt["cells"][ci]["regions"]["boxes_original"] == [cell["regions"]["boxes_original"][i] for i in kidx]
                kept.append(kidx)
            cases.append({"width": w, "height": h, "rows": rows, "cols": cols, "overlap": ov, "threshold": thr,
                          "cell_coordinates": [c["cell_coordinates"] for c in grid_info["cells"]],
                          "boxes_original": [c["regions"]["boxes_original"] for c in grid_info["cells"]],
                          "kept": kept})
    return cases


_R1 = None


def r1_cells(w, h, rows, cols, ov):
    """Cell coordinate dicts exactly as the reference emits them (mixed int/float)."""
    import cv2
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "p.png")
        cv2.imwrite(path, np.zeros((h, w), np.uint8))
        return [c["coordinates"] for c in _R1.split_image_into_grid(path, rows, cols, ov)]


def pooled_case(w, h, rows, cols, n, seed, thr_edge=10):
    det = synth.page_detections(w, h, rows, cols, 20.0, n, seed)
    cells = det["cells"]
    bp = det["boxes_local"] + cells[det["box_cell"]][:, [0, 1, 0, 1]]
    return bp, det["scores"], det["classes"]


def gen_stage3(r3):
    arrays = {}
    meta = []
    specs = [("cfg2_like", 3801, 5601, 2, 2, 2000, 11, 0.5), ("cfg3_small", 8000, 6000, 4, 4, 3000, 12, 0.5),
             ("thr03", 4000, 5443, 3, 3, 1200, 13, 0.3), ("thr07", 2778, 4187, 2, 2, 900, 14, 0.7),
             ("thr0", 3000, 3000, 2, 2, 400, 15, 0.0)]
    for name, w, h, rows, cols, n, seed, thr in specs:
        b, s, c = pooled_case(w, h, rows, cols, n, seed)
        arrays[name] = (b, s, c, thr)
    # adversarial: score ties, duplicate boxes, degenerate (zero-area) boxes, touching boxes
    rng = np.random.default_rng(99)
    b, s, c = pooled_case(2000, 2000, 2, 2, 300, 16)
    s = np.round(s, 1)                       # many exact score ties
    b[50:60] = b[40:50]                      # exact duplicates
    c[50:60] = c[40:50]
    b[100:105, 2] = b[100:105, 0]            # zero width
    b[105:110, 3] = b[105:110, 1]            # zero height
    b[120] = [10, 10, 20, 20]
    b[121] = [20, 10, 30, 20]                # touching -> IoU 0
    c[120] = c[121] = 1.0
    arrays["adversarial"] = (b, s, c, 0.5)
    arrays["single"] = (np.array([[1.0, 2.0, 3.0, 4.0]]), np.array([0.5]), np.array([1.0]), 0.5)
    out = {}
    for name, (b, s, c, thr) in arrays.items():
        names = synth.class_names_of(c)
        fb, fs, fc, fn = r3.apply_non_max_suppression(b.tolist(), s.tolist(), c.tolist(), names, thr)
        # recover positions: replay with an index list riding along as "class_names"
        _, _, _, fidx = r3.apply_non_max_suppression(b.tolist(), s.tolist(), c.tolist(), list(range(len(s))), thr)
        assert [b[i].tolist() for i in fidx] == fb and [s[i] for i in fidx] == fs
        out[name + "_boxes"] = b
        out[name + "_scores"] = s
        out[name + "_classes"] = c
        out[name + "_thr"] = np.array([thr])
        out[name + "_kept"] = np.array(fidx, np.int32)
        meta.append(name)
    assert r3.apply_non_max_suppression([], [], [], [], 0.5) == ([], [], [], [])
    # calculate_iou known answers
    pairs = np.concatenate([rng.uniform(0, 100, (200, 8)), rng.integers(0, 12, (200, 8)).astype(np.float64)])
    pairs[:, 2:4] = np.maximum(pairs[:, 2:4], pairs[:, 0:2])
    pairs[:, 6:8] = np.maximum(pairs[:, 6:8], pairs[:, 4:6])
    out["iou_pairs"] = pairs
    out["iou_values"] = np.array([r3.calculate_iou(p[:4].tolist(), p[4:].tolist()) for p in pairs])
    out["cases"] = np.array(meta)
    return out


def gen_stage45_synth(r3, r4, r5):
    out = []
    specs = [(8000, 6000, 4, 4, 10000, 21, 0.2, 0.3), (3801, 5601, 2, 2, 2000, 22, 0.2, 0.3),
             (2778, 4187, 2, 2, 1500, 23, 0.2, 0.3), (7934, 5755, 3, 3, 4000, 24, 0.05, 0.5),
             (4000, 5443, 2, 2, 2000, 25, 1.0, 0.1), (3000, 3000, 2, 2, 300, 26, -0.1, 0.3),
             (2000, 2000, 2, 2, 100, 27, 0.2, 0.995)]
    from oracle.nms_fast import nms_pick_order_c
    for w, h, rows, cols, n, seed, mm, mc in specs:
        b, s, c = pooled_case(w, h, rows, cols, n, seed)
        k = nms_pick_order_c(b, s, c, 0.5)
        b, s, c = b[k], s[k], c[k]
        names = synth.class_names_of(c)
        widths = [bb[2] - bb[0] for bb, nn in zip(b.tolist(), names) if nn == "plain_text"]
        bins = r4.bin_widths(widths, mm, w)
        med = r4.calculate_median_width(bins)
        centers, cw = ([], [])
        if med > 0:
            centers, cw = r5.find_column_centers(b.tolist(), names, s.tolist(), w, h, med, mc)
        out.append({"width": w, "height": h, "rows": rows, "cols": cols, "n": n, "seed": seed,
                    "min_margin_percent": mm, "min_confidence": mc,
                    "input_sha256": hashlib.sha256(b.tobytes() + s.tobytes() + c.tobytes()).hexdigest(),
                    "n_kept": int(len(k)), "n_bins": len(bins), "median_width": float(med),
                    "column_centers": [float(x) for x in centers], "column_widths": [float(x) for x in cw]})
    return out


def gen_tile_nms():
    """torchvision.ops.nms exactly as 1_doclayout_bboxes.py:219-223 calls it (float32 tensors, CPU),
    on per-tile synthetic detections at the reference's default 0.45 and two other thresholds."""
    import torch
    import torchvision
    out, names = {}, []
    specs = [("t045", 2800, 2100, 1200, 61, 0.45), ("t030", 1024, 1024, 700, 62, 0.3), ("t070", 2000, 1500, 900, 63, 0.7),
             ("ties", 1500, 1500, 400, 64, 0.45), ("dense", 800, 800, 1500, 65, 0.45)]
    for name, w, h, n, seed, thr in specs:
        d = synth.page_detections(w, h, 1, 1, 20.0, n, seed, dups=3)
        b = d["boxes_local"].astype(np.float32)
        s = d["scores"].astype(np.float32)
        if name == "ties":
            s = np.round(s, 1).astype(np.float32)
            b[30:40] = b[20:30]
        keep = torchvision.ops.nms(boxes=torch.tensor(b), scores=torch.tensor(s), iou_threshold=thr).numpy()
        out[name + "_boxes"], out[name + "_scores"] = b, s
        out[name + "_thr"], out[name + "_keep"] = np.array([thr]), keep.astype(np.int64)
        names.append(name)
    out["cases"] = np.array(names)
    out["torchvision_version"] = np.array([torchvision.__version__])
    return out


def gen_cli_tree(mods):
    """Run the reference's own main()s for stages 2-5 on a synthetic stage-1 tree (tests/cli_tree.py)
    and record every JSON they write, with the tree root normalised."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cli_tree
    with tempfile.TemporaryDirectory() as root:
        cli_tree.build_stage1_tree(root)
        argv = cli_tree.stage_argv(root)
        old = sys.argv
        try:
            for stage, mod in zip((2, 3, 4, 5), mods):
                sys.argv = [f"stage{stage}"] + argv[stage]
                mod.main()
        finally:
            sys.argv = old
        return cli_tree.collect_outputs(root)


def main():
    global _R1
    os.makedirs(OUT, exist_ok=True)
    r2, r3, r4, r5 = (load_ref(s) for s in ("2_edge_box_filter", "3_combine_grids",
                                            "4_extract_median_widths", "5_detect_column_centers"))
    _R1 = load_ref("1_doclayout_bboxes")
    for m in (_R1, r2, r3, r4, r5):
        m.logger.setLevel("ERROR")
    pages = f1_pages()
    with gzip.open(os.path.join(OUT, "f1_pages.json.gz"), "wt") as f:
        json.dump(pages, f)
    # reference KAT: NMS on its own output is the identity
    for p in pages:
        fb, fs, fc, fn = r3.apply_non_max_suppression(p["boxes"], p["scores"], p["classes"], p["class_names"], 0.5)
        assert fb == p["boxes"] and fs == p["scores"], p["name"]
    with open(os.path.join(OUT, "f4_stage45.json"), "w") as f:
        json.dump(gen_f4(pages, r4, r5), f, indent=1)
    with open(os.path.join(OUT, "stage1_geometry.json"), "w") as f:
        json.dump(gen_stage1(_R1), f)
    with gzip.open(os.path.join(OUT, "stage2_filter.json.gz"), "wt") as f:
        json.dump(gen_stage2(r2), f)
    np.savez_compressed(os.path.join(OUT, "stage3_nms.npz"), **gen_stage3(r3))
    with open(os.path.join(OUT, "stage45_synth.json"), "w") as f:
        json.dump(gen_stage45_synth(r3, r4, r5), f, indent=1)
    np.savez_compressed(os.path.join(OUT, "stage1_tile_nms.npz"), **gen_tile_nms())
    with gzip.open(os.path.join(OUT, "cli_tree.json.gz"), "wt") as f:
        json.dump(gen_cli_tree((r2, r3, r4, r5)), f)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
