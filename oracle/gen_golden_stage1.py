"""Golden tree of the WHOLE command-line chain, made by the unmodified reference: stage 1's own `main()`
(1_doclayout_bboxes.py:682-785: process_image, process_image_with_grid, detect_regions with its
torchvision NMS, the tile PNG round trip) with a stub in place of the network, then 2_edge_box_filter.py
`--process_grids` (2:579-649: filter_edge_boxes on the per-cell files), then stages 3, 4 and 5 on that output.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).

    python -m oracle.gen_golden_stage1      # writes tests/golden/stage1_chain.json.gz

What is patched, and only this: `YOLODocumentLayoutDetector.__init__` (1:80-189 downloads weights and clones a
repository — no network here) is replaced by one that installs `tests/stub_detector.raw_detections` as
`self.model`; everything downstream of `self.model.predict(...)` is the reference's code.  Directory listings
(`glob.glob`, `os.listdir`) are sorted while stage 3 runs, because the reference takes them in filesystem order
(3:157,166,178,183) and the pooled order decides ties and `source_jsons`; the golden pins the sorted order.

The fixture holds, with the tree root replaced by <ROOT>: the text of every JSON file of every stage, the
sha256 of the decoded pixels of every tile image stage 1 wrote (1:568), and each tile's shape.
"""
from __future__ import annotations

import glob
import gzip
import json
import os
import sys
import tempfile
import types


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden")

import stub_detector  # noqa: E402
from oracle.gen_golden import load_ref  # noqa: E402

from stage1_chain import STAGES, chain_argv, collect, write_pages  # noqa: E402


class _StubModel:
    """What `from doclayout_yolo import YOLOv10; YOLOv10(path)` would have been (1:178-179), as far as
    detect_regions touches it (1:205-215)."""

    def predict(self, image, imgsz=None, conf=None, device=None):
        import torch
        b, c, s = stub_detector.raw_detections(image.filename, image.width, image.height)
        boxes = types.SimpleNamespace(xyxy=torch.from_numpy(b), cls=torch.from_numpy(c), conf=torch.from_numpy(s))
        return [types.SimpleNamespace(boxes=boxes)]


def run_reference_chain(root: str):
    r1, r2, r3, r4, r5 = (load_ref(s) for s in ("1_doclayout_bboxes", "2_edge_box_filter", "3_combine_grids",
                                                 "4_extract_median_widths", "5_detect_column_centers"))

    def stub_init(self, conf_threshold=0.1, iou_threshold=0.45, device=None, **_):
        self.device, self.conf_threshold, self.iou_threshold, self.image_size = device or "cpu", conf_threshold, iou_threshold, 1024
        self.model = _StubModel()

    r1.YOLODocumentLayoutDetector.__init__ = stub_init
    for m in (r1, r2, r3, r4, r5):
        m.logger.setLevel("ERROR")
    argv = chain_argv(root)
    old_argv, old_glob, old_listdir = sys.argv, glob.glob, os.listdir
    try:
        for stage, mod in ((1, r1), (2, r2), (3, r3), (4, r4), (5, r5)):
            sys.argv = [f"stage{stage}"] + argv[stage]
            if stage == 3:
                glob.glob = lambda *a, **k: sorted(old_glob(*a, **k))
                os.listdir = lambda *a, **k: sorted(old_listdir(*a, **k))
            mod.main()
            glob.glob, os.listdir = old_glob, old_listdir
    finally:
        sys.argv, glob.glob, os.listdir = old_argv, old_glob, old_listdir


def main():
    with tempfile.TemporaryDirectory() as root:
        write_pages(root)
        run_reference_chain(root)
        out = collect(root)
    n2 = sum(1 for k in out["files"] if k.startswith(STAGES[1] + os.sep + "grid_"))
    assert n2 > 0 and out["tiles"], "the reference wrote no per-cell files?"
    with gzip.open(os.path.join(OUT, "stage1_chain.json.gz"), "wt") as f:
        json.dump(out, f)
    print(f"stage1_chain.json.gz: {len(out['files'])} JSON files ({n2} per-cell stage-2 files), {len(out['tiles'])} tile images")


if __name__ == "__main__":
    main()
