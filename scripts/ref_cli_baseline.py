"""CLI-level CPU baseline: the UNMODIFIED reference `main()`s of stages 2-5, timed on a
synthetic directory tree — wall time per stage, including their JSON, image codec and visualisation work.

Build container only (imports /root/reference; no GPU needed).  Prints one JSON line per stage; the tree is the one
scripts/bench_cli_stages.py times this repository's command lines on (8000x6000 pages, full page + 2x2 + 3x3 + 4x4
documents, `--boxes` detections per page), plus the page images the reference insists on opening.

    python scripts/ref_cli_baseline.py [--pages 4] [--boxes 10000]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import cv2  # noqa: E402
import numpy as np  # noqa: E402

from bench_cli_stages import write_tree  # noqa: E402
from oracle.gen_golden import load_ref  # noqa: E402
from multimodal_embeddings_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=4)
    ap.add_argument("--boxes", type=int, default=10000)
    a = ap.parse_args()
    mods = {2: load_ref("2_edge_box_filter"), 3: load_ref("3_combine_grids"), 4: load_ref("4_extract_median_widths"),
            5: load_ref("5_detect_column_centers")}
    for m in mods.values():
        m.logger.setLevel("ERROR")
    with tempfile.TemporaryDirectory() as tmp:
        s1 = os.path.join(tmp, "1_doclayout_parsed")
        write_tree(s1, a.pages, a.boxes)
        # the documents name /corpus/page_NNNN.png: point them at real files (the reference opens them, 2:189-203,
        # 4:272, 5:381, and draws its visualisations on them)
        img_dir = os.path.join(tmp, "0_oriented_images")
        os.makedirs(img_dir)
        page = synth.newspaper_page(8000, 6000, 1)
        for p in range(a.pages):
            cv2.imwrite(os.path.join(img_dir, f"page_{p:04d}.png"), page)
        for fn in os.listdir(os.path.join(s1, "json")):
            path = os.path.join(s1, "json", fn)
            with open(path) as f:
                text = f.read().replace("/corpus/page_", img_dir + "/page_")
            with open(path, "w") as f:
                f.write(text)
        r = lambda *x: os.path.join(tmp, *x)  # noqa: E731
        argv = {2: ["--input_folder", s1, "--output_folder", r("2")],
                3: ["--input_folder", r("2"), "--output_folder", r("3")],
                4: ["--input_folder", r("3", "json"), "--output_folder", r("4")],
                5: ["--input_folder", r("3", "json"), "--median_folder", r("4", "json"), "--output_folder", r("5")]}
        old = sys.argv
        for stage in (2, 3, 4, 5):
            sys.argv = [f"stage{stage}"] + argv[stage]
            t0 = time.perf_counter()
            mods[stage].main()
            dt = time.perf_counter() - t0
            out = argv[stage][argv[stage].index("--output_folder") + 1]
            n_json = sum(len([f for f in fs if f.endswith(".json")]) for _, _, fs in os.walk(out))
            print(json.dumps({"reference_stage": stage, "pages": a.pages, "boxes_per_page_in": a.boxes, "seconds": round(dt, 2),
                              "seconds_per_page": round(dt / a.pages, 2), "json_files_written": n_json,
                              "cores": 1, "host": f"build container, {os.cpu_count()} vCPU"}), flush=True)
        sys.argv = old


if __name__ == "__main__":
    main()
