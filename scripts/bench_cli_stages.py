"""Wall time of the stage-3 CLI (3_combine_grids.py drop-in) on a synthetic stage-2 tree, with the device
reader + writer and with CPython's json on both sides (PG_PYTHON_JSON=1) — what the §8f rank-2 work buys the
drop-in command line.  Prints one JSON line.

    python scripts/bench_cli_stages.py [--pages 16] [--boxes 10000]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_embeddings_b200 import cli, synth  # noqa: E402


def write_tree(root, pages, boxes):
    """<root>/json/<base>.json (full page) and <base>_grid_RxC.json for 2x2, 3x3, 4x4, as stage 2 leaves them."""
    w, h = 8000, 6000
    os.makedirs(os.path.join(root, "json"), exist_ok=True)
    total = 0
    for p in range(pages):
        base = f"page_{p:04d}"
        for rows, cols in ((1, 1), (2, 2), (3, 3), (4, 4)):
            d = synth.page_detections(w, h, rows, cols, 20.0, boxes // 4, synth.PAGE_SEED0 + 31 * p + rows)
            bo = d["boxes_local"] + d["cells"][d["box_cell"]][:, [0, 1, 0, 1]]
            names = synth.class_names_of(d["classes"])
            if (rows, cols) == (1, 1):
                doc = {"image_path": f"/corpus/{base}.png", "image_size": {"width": w, "height": h},
                       "parameters": {"conf_threshold": 0.1, "iou_threshold": 0.45},
                       "boxes": bo.tolist(), "classes": d["classes"].tolist(), "scores": d["scores"].tolist(),
                       "class_names": names}
                path = os.path.join(root, "json", f"{base}.json")
            else:
                cells = []
                for ci in range(rows * cols):
                    m = d["box_cell"] == ci
                    c = d["cells"][ci]
                    cells.append({"cell_path": f"/corpus/grid_{rows}x{cols}/images/{base}_c{ci}.png",
                                  "cell_json_path": f"/corpus/grid_{rows}x{cols}/json/{base}_c{ci}.json",
                                  "cell_coordinates": {"x_start": float(c[0]), "y_start": float(c[1]), "x_end": float(c[2]),
                                                       "y_end": float(c[3])},
                                  "row": ci // cols + 1, "col": ci % cols + 1,
                                  "regions": {"boxes": d["boxes_local"][m].tolist(), "boxes_original": bo[m].tolist(),
                                              "classes": d["classes"][m].tolist(), "scores": d["scores"][m].tolist(),
                                              "class_names": [n for n, k in zip(names, m) if k]}})
                doc = {"original_image_path": f"/corpus/{base}.png", "cells": cells,
                       "grid_config": {"rows": rows, "cols": cols, "overlap_percentage": 20.0}}
                path = os.path.join(root, "json", f"{base}_grid_{rows}x{cols}.json")
            with open(path, "w") as f:
                json.dump(doc, f, indent=2)
            total += os.path.getsize(path)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=16)
    ap.add_argument("--boxes", type=int, default=10000)
    ap.add_argument("--stage", type=int, default=3, choices=[2, 3, 4, 5])
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        s2 = os.path.join(tmp, "2_edge_box_filtered")
        in_bytes = write_tree(s2, a.pages, a.boxes)
        s3, s4 = os.path.join(tmp, "3_combined_bboxes"), os.path.join(tmp, "4_medians_extracted")
        if a.stage >= 4:  # the later stages read what the earlier ones wrote
            assert cli.main_stage3(["--input_folder", s2, "--output_folder", s3]) == 0
            in_bytes = sum(os.path.getsize(os.path.join(s3, "json", f)) for f in os.listdir(os.path.join(s3, "json")))
        if a.stage == 5:
            assert cli.main_stage4(["--input_folder", os.path.join(s3, "json"), "--output_folder", s4, "--no_image_check"]) == 0
        main_fn = {2: cli.main_stage2, 3: cli.main_stage3, 4: cli.main_stage4, 5: cli.main_stage5}[a.stage]
        argv = {2: ["--input_folder", s2, "--no_image_check"], 3: ["--input_folder", s2],
                4: ["--input_folder", os.path.join(s3, "json"), "--no_image_check"],
                5: ["--input_folder", os.path.join(s3, "json"), "--median_folder", os.path.join(s4, "json"),
                    "--no_image_check"]}[a.stage]
        files_per_page = {2: 4, 3: 1, 4: 1, 5: 1}[a.stage]
        times, outs = {}, {}
        for mode in ("warmup", "device", "cpython"):
            out = os.path.join(tmp, f"out_{mode}")
            if mode == "cpython":
                os.environ["PG_PYTHON_JSON"] = "1"
            t0 = time.perf_counter()
            assert main_fn(argv + ["--output_folder", out]) == 0
            times[mode] = time.perf_counter() - t0
            os.environ.pop("PG_PYTHON_JSON", None)
            outs[mode] = {f: open(os.path.join(out, "json", f), "rb").read() for f in sorted(os.listdir(os.path.join(out, "json")))}
        assert outs["device"] == outs["cpython"] and len(outs["device"]) == a.pages * files_per_page
        out_bytes = sum(len(v) for v in outs["device"].values())
        what = {2: "stage-2 CLI wall time (read 4 stage-1 files per page, edge filter, write 4 files)",
                3: "stage-3 CLI wall time (read 4 stage-2 files per page, merge, write the record)",
                4: "stage-4 CLI wall time (read the stage-3 record, width median, write a small file)",
                5: "stage-5 CLI wall time (read the stage-3 record and the median, column peaks, write a small file)"}[a.stage]
        print(json.dumps({"what": what,
                          "pages": a.pages, "boxes_per_page_in": a.boxes, "input_json_mb": in_bytes / 1e6,
                          "output_json_mb": out_bytes / 1e6, "seconds_device_json": times["device"],
                          "seconds_cpython_json": times["cpython"], "pages_per_s_device_json": a.pages / times["device"],
                          "pages_per_s_cpython_json": a.pages / times["cpython"],
                          "speedup": times["cpython"] / times["device"], "outputs_identical": True}))


if __name__ == "__main__":
    main()
