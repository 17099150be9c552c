"""Inputs, command lines and output collection of the stage-1 -> stage-5 chain test, shared by the golden
generator (oracle/gen_golden_stage1.py: the unmodified reference with a stub network) and
tests/test_gpu_cli.py (this repository's command lines with the same stub as a detector plug-in)."""
import hashlib
import os

import numpy as np

import stub_detector
from multimodal_embeddings_b200 import synth

STAGES = ["1_doclayout_parsed", "2_edge_box_filtered", "3_combined_bboxes", "4_medians_extracted", "5_column_detection"]


def write_pages(root: str) -> str:
    """The input scans: seeded synthetic pages, written losslessly (PNG / BMP)."""
    import cv2
    folder = os.path.join(root, "0_oriented_images")
    os.makedirs(folder, exist_ok=True)
    for name, w, h, seed in stub_detector.PAGES:
        assert cv2.imwrite(os.path.join(folder, name), synth.page_pixels(w, h, seed))
    return folder


def chain_argv(root: str) -> dict:
    r = lambda *p: os.path.join(root, *p)  # noqa: E731
    return {
        1: ["--input_folder", r("0_oriented_images"), "--output_folder", r(STAGES[0]), "--grids", stub_detector.GRIDS,
            "--overlap", str(stub_detector.OVERLAP)],
        2: ["--input_folder", r(STAGES[0]), "--output_folder", r(STAGES[1]), "--process_grids"],
        3: ["--input_folder", r(STAGES[1]), "--output_folder", r(STAGES[2])],
        4: ["--input_folder", r(STAGES[2], "json"), "--output_folder", r(STAGES[3])],
        5: ["--input_folder", r(STAGES[2], "json"), "--median_folder", r(STAGES[3], "json"), "--output_folder", r(STAGES[4])],
    }


def collect(root: str) -> dict:
    """relative path -> file text (root replaced) for every .json under the stage folders; tile images ->
    {"sha256", "shape"} of the decoded array."""
    import cv2
    files, tiles = {}, {}
    for stage in STAGES:
        base = os.path.join(root, stage)
        for d, _, fs in os.walk(base):
            for fn in sorted(fs):
                path = os.path.join(d, fn)
                rel = os.path.relpath(path, root)
                if fn.endswith(".json"):
                    with open(path) as f:
                        files[rel] = f.read().replace(root, "<ROOT>")
                elif os.sep + "images" + os.sep in path:
                    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
                    tiles[rel] = {"sha256": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest(),
                                  "shape": list(img.shape)}
    return {"files": files, "tiles": tiles}
