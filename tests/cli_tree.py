"""Synthetic stage-1 output tree (the layout 1_doclayout_bboxes.py writes) used to compare the
reference CLIs (run in the build container by oracle/gen_golden.py) with this repo's CLIs."""
import json
import os

import numpy as np

from multimodal_embeddings_b200 import synth
from oracle import tiler as ot

PAGES = [("Gazette 1901 page_0001", 3801, 5601, 501, 900), ("Herald 1888 page_0002", 2778, 4187, 502, 600),
         ("Empty page_0003", 1200, 900, 503, 0)]
ROWS, COLS, OVERLAP = 2, 2, 20.0


def build_stage1_tree(root: str) -> str:
    from PIL import Image
    img_dir = os.path.join(root, "0_oriented_images")
    out = os.path.join(root, "1_doclayout_parsed")
    os.makedirs(img_dir, exist_ok=True)
    os.makedirs(os.path.join(out, "json"), exist_ok=True)
    params = {"conf_threshold": 0.1, "iou_threshold": 0.45}
    for name, w, h, seed, n in PAGES:
        image_path = os.path.join(img_dir, name + ".png")
        Image.fromarray(np.zeros((h, w), np.uint8)).save(image_path, compress_level=1)
        full = synth.page_detections(w, h, 1, 1, OVERLAP, max(n // 3, 1 if n else 0), seed)
        _dump({"image_path": image_path, "image_size": {"width": w, "height": h}, "parameters": params,
               "boxes": full["boxes_local"].tolist(), "classes": full["classes"].tolist(),
               "scores": full["scores"].tolist(), "class_names": synth.class_names_of(full["classes"])},
              os.path.join(out, "json", name + ".json"))
        det = synth.page_detections(w, h, ROWS, COLS, OVERLAP, n, seed + 1000)
        cells = ot.grid_cells(w, h, ROWS, COLS, OVERLAP)
        info = {"original_image_path": image_path,
                "grid_config": {"rows": ROWS, "cols": COLS, "overlap_percentage": OVERLAP}, "cells": []}
        for ci, cell in enumerate(cells):
            m = det["box_cell"] == ci
            local = det["boxes_local"][m].tolist()
            info["cells"].append({
                "cell_path": os.path.join(out, f"grid_{ROWS}x{COLS}", "images", f"{name}_row{cell['row']}_col{cell['col']}.png"),
                "cell_json_path": os.path.join(out, f"grid_{ROWS}x{COLS}", "json", f"{name}_row{cell['row']}_col{cell['col']}.json"),
                "cell_coordinates": cell["coordinates"], "row": cell["row"], "col": cell["col"],
                "regions": {"boxes": local, "boxes_original": ot.translate_boxes(local, cell["coordinates"]),
                            "classes": det["classes"][m].tolist(), "scores": det["scores"][m].tolist(),
                            "class_names": synth.class_names_of(det["classes"][m])}})
        _dump(info, os.path.join(out, "json", f"{name}_grid_{ROWS}x{COLS}.json"))
    return out


def _dump(obj, path):
    with open(path, "w") as f:
        json.dump(obj, f, indent=2)


STAGE_DIRS = ["2_edge_box_filtered", "3_combined_bboxes", "4_medians_extracted", "5_column_detection"]


def stage_argv(root: str):
    r = lambda *p: os.path.join(root, *p)  # noqa: E731
    return {
        2: ["--input_folder", r("1_doclayout_parsed"), "--output_folder", r("2_edge_box_filtered")],
        3: ["--input_folder", r("2_edge_box_filtered"), "--output_folder", r("3_combined_bboxes")],
        4: ["--input_folder", r("3_combined_bboxes", "json"), "--output_folder", r("4_medians_extracted")],
        5: ["--input_folder", r("3_combined_bboxes", "json"), "--median_folder", r("4_medians_extracted", "json"),
            "--output_folder", r("5_column_detection")],
    }


def collect_outputs(root: str) -> dict:
    """relative path -> JSON text with the tree root replaced by <ROOT> (key order preserved)."""
    out = {}
    for d in STAGE_DIRS:
        jd = os.path.join(root, d, "json")
        if not os.path.isdir(jd):
            continue
        for fn in sorted(os.listdir(jd)):
            if fn.endswith(".json"):
                with open(os.path.join(jd, fn)) as f:
                    out[f"{d}/json/{fn}"] = json.dumps(json.load(f)).replace(root, "<ROOT>")
    return out
